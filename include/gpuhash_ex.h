/*
 * gpuhash_ex.h -- extended C ABI of the B200-native Mega-KV hash index.
 *
 * Plain C: pointers, sizes and integers only (streams travel as void*), so it can
 * be bound from C, cgo, ctypes ... without CUDA headers.  The three legacy entry
 * points of libgpuhash.h are thin wrappers over the *_ex calls below using the
 * process-wide default geometry.
 *
 * What each group replaces in the reference (pzrq/megakv):
 *   geometry            compile-time macros of libgpuhash/gpu_hash.h:46-76
 *   *_ex launches       gpu_hash_search / _insert / _delete, gpu_hash.cu:482-593
 *   gpuhash_index_*     the per-cycle transfer + launch loop of the scheduler,
 *                       src/mega_scheduler.c:392-504, and the pinned/device batch
 *                       buffers of src/mega_recv.c:120-213
 *   gpuhash_bench_*     the micro-benchmarks libgpuhash/test/back/ *_stream.c
 *   gpuhash_dev_* etc.  cudaMalloc/cudaMemcpy plumbing the reference's tests do inline
 *
 * All functions return 0 on success or a cudaError_t value (> 0); -1 = bad argument.
 */
#ifndef GPUHASH_EX_H
#define GPUHASH_EX_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPUHASH_CUCKOO   0u     /* HASH_CUCKOO   gpu_hash.h:73 */
#define GPUHASH_2CHOICE  1u     /* HASH_2CHOICE  gpu_hash.h:72 */

/* table layouts.  The caller of the legacy ABI never interprets table bytes (mega_scheduler.c:273-274 only
 * allocates and zero-fills), so the layout is the library's choice; all-zero == empty in both.
 *   PAIRS      slot l = 8-byte {sig, loc} at byte 8*l of the 64 B bucket: (sig, loc) changes are one 64-bit CAS
 *   REFERENCE  bucket_t of gpu_hash.h:79-82 (sig[8] then loc[8]); for callers that memcpy tables in that layout.
 *              A commit is then two steps -- CAS on the signature word, store / exchange of the location word -- the
 *              reference's own publication window: two cuckoo evictions that hit one slot back to back (launches of
 *              several streams, or one launch at high load) can each carry the other's location on, leaving a
 *              (sig, loc) mismatch.  The reference excludes that by running one CUDA block per insert partition; this
 *              library runs grid-wide.  Use PAIRS for any table that takes concurrent cuckoo inserts at load factors
 *              where evictions happen; REFERENCE is exact for searches, deletes, two-choice and low-load inserts. */
#define GPUHASH_LAYOUT_PAIRS      0u
#define GPUHASH_LAYOUT_REFERENCE  1u

/* insert flags */
#define GPUHASH_INSERT_SERIAL  1u   /* one thread, batch order: slot-for-slot equal to a sequential run */

typedef struct gpuhash_geom_s {
	uint32_t hash_mask;     /* buckets of this table - 1 (a shard holds a slice of the logical table) */
	uint32_t block_mask;    /* BLOCK_HASH_MASK of the LOGICAL table: low bits the alternate bucket may change */
	uint32_t algo;          /* GPUHASH_CUCKOO | GPUHASH_2CHOICE */
	uint32_t max_cuckoo;    /* MAX_CUCKOO_NUM (5) */
	uint32_t layout;        /* GPUHASH_LAYOUT_PAIRS (set by gpuhash_geom_init) | GPUHASH_LAYOUT_REFERENCE */
} gpuhash_geom_t;

typedef struct gpuhash_stats_s {
	unsigned long long ins_skipped, ins_updated, ins_placed_b1, ins_placed_b2, ins_to_b2,
	                   ins_displaced, ins_dropped, ins_overwritten, ins_cas_retry, ins_gave_up,
	                   chain_hist[8],
	                   del_zeroed, del_requests_hit,
	                   search_hits_b1, search_hits_b2;
} gpuhash_stats_t;

/* ---- geometry ---- */
int    gpuhash_geom_init(gpuhash_geom_t *g, int mem_p, unsigned algo);
/* shard `log2_shards` of a logical 2^mem_p_total-byte table: local table is 2^(mem_p_total-log2_shards) bytes */
int    gpuhash_geom_init_shard(gpuhash_geom_t *g, int mem_p_total, int log2_shards, unsigned algo);
/* gpuhash_geom_init + the layout that suits the table on the current device: PAIRS, except REFERENCE for a two-choice table
 * that fits in L2 (cudaDevAttrL2CacheSize), where L2 sectors -- not DRAM line fills -- are the cost of a probe */
int    gpuhash_geom_init_auto(gpuhash_geom_t *g, int mem_p, unsigned algo);
size_t gpuhash_table_bytes(const gpuhash_geom_t *g);
/* in-place rewrite of a table from g->layout to to_layout (async on stream); the caller then sets g->layout */
int    gpuhash_table_convert(const gpuhash_geom_t *g, void *table_d, unsigned to_layout, void *stream);
void   gpuhash_set_default_geom(const gpuhash_geom_t *g);   /* what the 3 legacy entry points use */
void   gpuhash_get_default_geom(gpuhash_geom_t *g);

/* ---- launch tuning (process-wide; defaults are what bench.py measures) ---- */
typedef struct gpuhash_tune_s {
	int search_qpt;          /* 0 = choose (default); -4 = four lanes per request, one L2 request per bucket;
	                            -5 = the same with the request/result batches staged through shared memory by
	                                 512 B bulk copies (TMA 1-D, mbarrier);
	                            -6 = four lanes per request, every warp on its own: 512 B tile in by one vector access,
	                                 four table loads per lane in flight, 512 B tile out (no shared memory, no barrier)
	                                 -- what 0 chooses for the pair layout and for tables beyond L2;
	                            1, 2, 4 = one thread per request, that many requests per thread;
	                            -1 = 4 lanes x 128-bit loads (reference layout only; comparison kernel) */
	int search_split_mode;   /* REFERENCE layout only: 0 = by table size, 1 = location word on hit only, 2 = whole buckets */
	int insert_ctas_per_sm;  /* grid of the count-independent insert kernel */
	int fused_cycle;         /* 1 (default): gpuhash_index_submit and the bench loops issue ONE launch per batch
	                            (gpuhash_cycle_ex); 0: one launch per operation kind, like the reference */
} gpuhash_tune_t;
void   gpuhash_set_tuning(const gpuhash_tune_t *t);
void   gpuhash_get_tuning(gpuhash_tune_t *t);

/* ---- asynchronous launches on device pointers (stream: cudaStream_t as void*, NULL = default) ---- */
int gpuhash_search_ex(const gpuhash_geom_t *g, const void *selem_d, void *out_d, const void *table_d,
		size_t n, gpuhash_stats_t *stats_d, void *stream);
/* The steps either side of the path, on the device:
 *   search_compact  one word per request -- bucket-1 hit, else bucket-2 hit, else 0: the sender's choice
 *                   (src/mega_send.c:411-414) made before the results cross the host link (4 B instead of 8 B per search)
 *   fold_keys       key bytes -> selem_t (src/mega_recv.c:349-362; fold != 0: the -DSIGNATURE XOR fold) */
int gpuhash_search_compact_ex(const gpuhash_geom_t *g, const void *selem_d, void *out32_d, const void *table_d,
		size_t n, gpuhash_stats_t *stats_d, void *stream);
int gpuhash_fold_keys_ex(const void *keys_d, size_t stride, unsigned nkey, int fold, size_t n, void *selem_out_d, void *stream);
int gpuhash_insert_ex(const gpuhash_geom_t *g, void *table_d, const void *const *blk_input_d,
		const int *blk_elem_num_d, int num_blks, gpuhash_stats_t *stats_d, unsigned flags, void *stream);
int gpuhash_insert_flat_ex(const gpuhash_geom_t *g, void *table_d, const void *ielem_d, size_t n,
		gpuhash_stats_t *stats_d, unsigned flags, void *stream);
int gpuhash_delete_ex(const gpuhash_geom_t *g, const void *delem_d, void *table_d, size_t n,
		gpuhash_stats_t *stats_d, unsigned flags, void *stream);

/* Allocates the library's small per-device state now instead of at first use (needed before capturing launches into a
 * CUDA graph; gpuhash_index_create and the bench loops call it).  Idempotent, current device. */
int gpuhash_init_device(void);

/* ---- one launch per scheduler cycle ----
 * The reference's cycle walks every worker's batch -- copies, gpu_hash_search, then gpu_hash_delete and gpu_hash_insert
 * on the worker's stream -- and synchronises once (src/mega_scheduler.c:393-420, 440-502, 504).  gpuhash_cycle_multi_ex is
 * that whole cycle as ONE kernel over a table of batch descriptors: per batch search -> delete -> insert (the reference's
 * in-stream order), batches unordered against each other (its streams).  Tiles of 64 requests are handed to warps by an
 * atomic ticket in phase-major order, so the kernel makes no assumption about CTA dispatch order or residency.
 *   batches_d     the descriptor table where the DEVICE reads it (device memory: every CTA reads it)
 *   batches_h     the same table on the host, only used to size the grid (NULL: full persistent grid)
 *   workspace_d   gpuhash_cycle_workspace_bytes(num_batches) bytes, zeroed once by the caller; a launch leaves them zero,
 *                 so one workspace serves one launch at a time (launches of one stream, or replays of one graph).
 *                 Word 2 is a sticky error flag: a phase wait that exceeded GPUHASH_CYCLE_TIMEOUT_MS (default 10 s).
 *   compact       != 0: one result word per search (the sender's choice, src/mega_send.c:411-414) instead of two
 * Pointers inside the descriptors are device-visible: device memory or pinned host memory (zero-copy). */
typedef struct gpuhash_batch_s {
	const void *search_in;  void *search_out;      /* selem_t[n_search]; loc_t[2 * n_search] (compact: loc_t[n_search]) */
	const void *delete_in;  const void *insert_in; /* delem_t[n_delete]; ielem_t[n_insert] */
	uint32_t n_search, n_delete, n_insert, reserved;
} gpuhash_batch_t;
#define GPUHASH_MAX_BATCHES 128
size_t gpuhash_cycle_workspace_bytes(int max_batches);
/* 1 if a phase wait of a cycle kernel on the current device timed out since the last reset (read after synchronising) */
int gpuhash_cycle_error(int reset);
/* cap on the cycle kernel's persistent grid, CTAs (of 8 warps) per SM; 0 = fill the GPU (default).  Host-link-bound cycles
 * (zero-copy batches) are served by a fraction of the warps; a small grid lets consecutive cycles be resident together. */
void gpuhash_set_cycle_ctas_per_sm(int n);
int gpuhash_cycle_multi_ex(const gpuhash_geom_t *g, void *table_d, const gpuhash_batch_t *batches_h,
		const gpuhash_batch_t *batches_d, int num_batches, int compact, void *workspace_d,
		gpuhash_stats_t *stats_d, void *stream);

/* The cycle of ONE worker in one launch (also what the legacy gpu_delete_insert, libgpuhash.h:53-62, runs).  Inserts are
 * either a flat batch (ielem_d, n_insert) or segments with device-side counts (blk_input_d, blk_elem_num_d, num_blks; then
 * ielem_d = NULL, n_insert = 0).  Any part may be empty.  gpuhash_cycle_ex takes its workspace from a per-device pool
 * (4096 slots handed out round-robin by an atomic counter): fine for direct calls; launches that are captured into CUDA
 * graphs and replayed next to other launches should own theirs (gpuhash_cycle_ws_ex, workspace_d as above, 1 batch). */
int gpuhash_cycle_ex(const gpuhash_geom_t *g, void *table_d,
		const void *selem_d, size_t n_search, void *out_d,
		const void *delem_d, size_t n_delete,
		const void *ielem_d, size_t n_insert,
		const void *const *blk_input_d, const int *blk_elem_num_d, int num_blks,
		gpuhash_stats_t *stats_d, void *stream);
int gpuhash_cycle_ws_ex(const gpuhash_geom_t *g, void *table_d,
		const void *selem_d, size_t n_search, void *out_d,
		const void *delem_d, size_t n_delete,
		const void *ielem_d, size_t n_insert,
		const void *const *blk_input_d, const int *blk_elem_num_d, int num_blks,
		int compact, void *workspace_d, gpuhash_stats_t *stats_d, void *stream);

/* ---- device / pinned memory and stream plumbing ---- */
int   gpuhash_device_count(void);
int   gpuhash_set_device(int dev);
int   gpuhash_device_info(int dev, int *sm_count, int *l2_bytes, size_t *free_bytes, size_t *total_bytes);
void *gpuhash_dev_alloc(size_t bytes);                 /* NULL on failure */
int   gpuhash_dev_free(void *p);
int   gpuhash_dev_memset(void *p, int v, size_t bytes, void *stream);
int   gpuhash_h2d(void *dst_d, const void *src_h, size_t bytes, void *stream);   /* async when src is pinned */
int   gpuhash_d2h(void *dst_h, const void *src_d, size_t bytes, void *stream);
void *gpuhash_host_alloc(size_t bytes);                /* pinned */
int   gpuhash_host_free(void *p);
void *gpuhash_stream_create(void);
int   gpuhash_stream_destroy(void *stream);
int   gpuhash_stream_sync(void *stream);
int   gpuhash_device_sync(void);
void *gpuhash_event_create(void);
int   gpuhash_event_destroy(void *ev);
int   gpuhash_event_record(void *ev, void *stream);
int   gpuhash_event_elapsed_ms(void *start, void *stop, float *ms);   /* synchronises on `stop` */
int   gpuhash_set_l2_fetch_granularity(int bytes);   /* cudaLimitMaxL2FetchGranularity of the current device */
int   gpuhash_get_l2_fetch_granularity(void);
const char *gpuhash_error_string(int err);
const char *gpuhash_build_info(void);

/* ---- roofline probe: n random 32 B sectors read from `table_d` (bytes must be a power of two) ----
 * mode 0: signature sector of a random 64 B bucket; 1: any random 32 B sector; 2: whole random 64 B bucket.
 * Writes the kernel time of the best of `iters` launches. */
int gpuhash_roofline_gather(const void *table_d, size_t table_bytes, size_t n, int mode,
		int loads_per_thread, int iters, float *best_ms, void *stream);

/* ---- the scheduler's device side: an index that owns its table, streams and staging ---- */
typedef struct gpuhash_index_s gpuhash_index_t;

/* workers = number of independent batch slots (reference: one per CPU receiver, <= 16, each with its
 * own stream, mega_scheduler.c:276-280); capacities are per worker per cycle (mega.c:136,143-144). */
gpuhash_index_t *gpuhash_index_create(int mem_p, unsigned algo, int workers,
		size_t max_search, size_t max_insert, size_t max_delete);
gpuhash_index_t *gpuhash_index_create_layout(int mem_p, unsigned algo, unsigned layout, int workers,
		size_t max_search, size_t max_insert, size_t max_delete);
void   gpuhash_index_destroy(gpuhash_index_t *ix);
void  *gpuhash_index_table(gpuhash_index_t *ix);                       /* device pointer */
const gpuhash_geom_t *gpuhash_index_geom(const gpuhash_index_t *ix);
void  *gpuhash_index_stream(gpuhash_index_t *ix, int worker);
int    gpuhash_index_clear(gpuhash_index_t *ix);
int    gpuhash_index_load(gpuhash_index_t *ix, const void *table_h);   /* host image: reference byte layout, always */
int    gpuhash_index_dump(gpuhash_index_t *ix, void *table_h);         /* (converted to/from the device layout)     */
int    gpuhash_index_stats(gpuhash_index_t *ix, gpuhash_stats_t *out, int reset);
int    gpuhash_index_enable_stats(gpuhash_index_t *ix, int on);
/* on: gpuhash_index_submit's host buffers are PINNED and the kernels access them directly over PCIe (no staging copies) */
int    gpuhash_index_set_zero_copy(gpuhash_index_t *ix, int on);
/* on: search_out_h receives ONE word per request -- the first non-zero of the reference's {out[2i], out[2i+1]}, i.e. the
 * sender's choice (mega_send.c:411-414) made on the device; halves the result bytes over the host link */
int    gpuhash_index_set_compact_results(gpuhash_index_t *ix, int on);

/* One scheduler cycle for one worker with HOST buffers (pinned or pageable), in the reference's order
 * search -> delete -> insert on the worker's stream (mega_scheduler.c:392-502).  Asynchronous: results
 * are in search_out_h after gpuhash_index_sync().  Any of the three parts may be empty. */
int gpuhash_index_submit(gpuhash_index_t *ix, int worker,
		const void *search_in_h, size_t n_search, void *search_out_h,
		const void *delete_in_h, size_t n_delete,
		const void *insert_in_h, size_t n_insert);
int gpuhash_index_sync(gpuhash_index_t *ix);                            /* mega_scheduler.c:504 */

/* One scheduler cycle for ALL workers in one call and ONE kernel launch (gpuhash_cycle_multi_ex): batches_h[w] holds
 * worker w's HOST buffers and counts, as gpuhash_index_submit takes them one worker at a time.  Zero-copy mode: the
 * buffers are pinned and the kernel reads/writes them itself; else they are staged through device copies owned by the
 * cycle's slot (n_* within the capacities given to gpuhash_index_create), so that staged cycles in flight overlap (copy-in
 * of one, kernel of the next, copy-out of a third), and host arrays that are ADJACENT in memory -- worker w+1's array
 * starting where worker w's ends, as when all batch buffers are carved out of one pinned block -- are moved by ONE copy
 * per array and direction (multi-megabyte copies run at 48-55 GB/s per direction, 0.5 MB ones at 24).  Asynchronous: returns a ticket >= 0 (or a negative
 * error); gpuhash_index_wait(ticket) returns once that cycle's results are in the host buffers -- up to
 * GPUHASH_INDEX_SLOTS cycles may be in flight, the way the reference's triple-buffered batches allow
 * (src/include/mega_batch.h:74-82); gpuhash_index_sync waits for everything.  Both return 0, a CUDA error, or -3 when a
 * phase wait inside a cycle kernel timed out (the batch is suspect). */
#define GPUHASH_INDEX_SLOTS 4
int gpuhash_index_submit_all(gpuhash_index_t *ix, const gpuhash_batch_t *batches_h, int num_batches);
/* Cycles in flight keep the reference's order by default: the kernel of cycle k+1 starts when the kernel of cycle k has
 * finished (the reference synchronises the device once per cycle, mega_scheduler.c:504), so a GET submitted one cycle after
 * a SET sees it; only the copies of neighbouring cycles (staged mode) and the host's submission run ahead.  on != 0 drops
 * that edge: consecutive cycle kernels may overlap (tail of one, head of the next: ~4 % more throughput on resident
 * batches) and requests of DIFFERENT cycles in flight are unordered against each other, like workers inside a cycle. */
int gpuhash_index_set_unordered_cycles(gpuhash_index_t *ix, int on);
int gpuhash_index_wait(gpuhash_index_t *ix, int ticket);

/* ---- the scheduler cycle without launches (megakv_b200/csrc/gpuhash_ring.cu; north_star (c)) ----
 * `rings` descriptor rings of `slots` entries in pinned host memory feed ONE persistent kernel (ctas_per_sm CTAs per SM,
 * default 4, split evenly over the rings).  gpuhash_ring_submit has the arguments of gpuhash_index_submit; the buffers
 * must be PINNED (cudaHostAlloc / cudaHostRegister): the kernel reads the requests from them and writes the results into
 * them itself.  It costs the host one 64-byte descriptor and no CUDA call; it returns the batch number to wait for (> 0)
 * or a negative error, and blocks only while the ring is full.  Batches of one ring run strictly in order, each as
 * search -> delete -> insert (the reference's per-stream order, mega_scheduler.c:392-502); rings are unordered against
 * each other.  The kernel parks itself after idle_ms without a doorbell (default 2000) and is relaunched on demand.
 * One ring object per device at a time; the calls of one ring object come from one host thread (the scheduler thread of
 * the reference, src/mega.c:410) -- there is no locking inside.
 * While the kernel is resident, anything that synchronises the whole DEVICE (cudaDeviceSynchronize, cudaFree, and so
 * gpuhash_index_sync / _stats / _dump / _clear and gpuhash_dev_free) blocks until the kernel parks itself (idle_ms):
 * call gpuhash_ring_park first.  Submit and wait notice a parked kernel, also one that parked with batches pending, and
 * relaunch it. */
typedef struct gpuhash_ring_s gpuhash_ring_t;
gpuhash_ring_t *gpuhash_ring_create(const gpuhash_geom_t *g, void *table_d, int rings, int slots, int ctas_per_sm, unsigned idle_ms);
long long gpuhash_ring_submit(gpuhash_ring_t *q, int ring,
		const void *search_in_h, size_t n_search, void *search_out_h,
		const void *delete_in_h, size_t n_delete, const void *insert_in_h, size_t n_insert);
int  gpuhash_ring_wait(gpuhash_ring_t *q, int ring, long long ticket, unsigned timeout_ms);   /* 0 ok, -2 timeout */
int  gpuhash_ring_drain(gpuhash_ring_t *q, unsigned timeout_ms);                               /* every ring, everything submitted */
int  gpuhash_ring_park(gpuhash_ring_t *q);                                                     /* stop the kernel (pending batches resume at the next submit) */
void gpuhash_ring_destroy(gpuhash_ring_t *q);
int  gpuhash_ring_ctas_per_ring(const gpuhash_ring_t *q);
int  gpuhash_ring_trace(gpuhash_ring_t *q, int ring, unsigned long long out8[8]);   /* device timeline of the latest batch (ns) */

/* ---- sharded index: routing kernels (megakv_b200/csrc/gpuhash_shard.cu; north_star (d)) ----
 * A logical table of 2^mem_p_total bytes is cut into G = 2^log2_shards contiguous bucket ranges; shard g holds
 * range g as a local table with gpuhash_geom_init_shard geometry.  The owner of a request is the top log2_shards
 * bits of (hash & hash_mask_total); both candidate buckets and every eviction target share it (gpu_hash.h:67-69).
 * Pointer arrays (dst_ptrs, seg_*_ptrs, peer_*_ptrs, staged_ptrs) are HOST arrays of G device pointers, which may
 * be local or peer (CUDA IPC) addresses.  Regions hold `cap` requests per (source, owner) pair.
 *   route_scatter    requests -> region d of dst_ptrs by owner d; counts_d[8] = how many went to each owner;
 *                    perm_d[d*cap + slot] = index of the request in `in` (NULL for insert/delete batches)
 *   route_publish    fused path: store my counts into every owner's inbox_count[my_rank], then flag[my_rank] = seq
 *   search_segments  look up num_seg regions (seg_count_d[s] requests each) and store 8 B results to seg_out_ptrs[s];
 *                    wait_seq != 0: first wait until flags_d[0..num_seg) >= wait_seq (2 s timeout -> *err_d = 1)
 *   results_publish  fused path: tell every origin its results are stored (flag[my_rank] = seq on each peer)
 *   route_gather     out[perm[d][j]] = staged[d][j]; optional flag wait like search_segments
 *   delete_segments  gpu_hash_delete semantics over regions (inserts use gpuhash_insert_ex: counts are ints) */
int gpuhash_route_scatter(const void *in_d, size_t n, int elem_words, uint32_t hash_mask_total, int log2_shards,
		const void *const *dst_ptrs, uint32_t *counts_d, uint32_t *perm_d, size_t cap, void *stream);
int gpuhash_route_publish(const uint32_t *counts_d, int log2_shards, int my_rank,
		const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs, uint32_t seq, void *stream);
int gpuhash_search_segments(const gpuhash_geom_t *g, const void *table_d, int num_seg,
		const void *const *seg_in_ptrs, const uint32_t *seg_count_d, const void *const *seg_out_ptrs,
		size_t max_total, const uint32_t *flags_d, uint32_t wait_seq, uint32_t *err_d, void *stream);
int gpuhash_results_publish(int log2_shards, int my_rank, const void *const *peer_flag_ptrs, uint32_t seq, void *stream);
int gpuhash_route_gather(const void *const *staged_ptrs, const uint32_t *perm_d, const uint32_t *counts_d,
		size_t cap, int log2_shards, void *out_d, size_t n, const uint32_t *flags_d, uint32_t wait_seq, uint32_t *err_d, void *stream);
int gpuhash_delete_segments(const gpuhash_geom_t *g, void *table_d, int num_seg, const void *const *seg_in_ptrs,
		const uint32_t *seg_count_d, size_t max_total, gpuhash_stats_t *stats_d, void *stream);
/* The same protocol with publication and waiting folded into the producing / consuming kernels (3 launches per routed
 * search, 2 per routed insert/delete):
 *   route_scatter_pub  waits until ack_flags_d[0..G) >= seq-1 (owners consumed the previous batch), scatters into dst_ptrs,
 *                      then the last CTA stores counts2_d[seq&1][d] to owner d's count cell and raises its flag to seq.
 *                      counts2_d = uint32[2][8], zero at first use; ticket_d = one zeroed uint32 per lane.
 *   serve              waits until req_flags_d[0..G) >= seq, then op 0: looks up the G regions and stores results to
 *                      seg_out_ptrs (the origins' staging regions); op 1 / 2: inserts / deletes them; the last CTA raises
 *                      the result flag (= consumption ack) to seq on every origin.
 *   route_gather       (above, with wait_seq = seq) brings the results into request order.
 * ack_flags_d / req_flags_d may be NULL: then the kernel does not wait and the caller orders the stream with
 * gpuhash_wait_flags (one tiny CTA) instead.  That is what megakv_b200/sharded.py does: a waiting CTA inside a large
 * kernel keeps an SM slot, and with many batches in flight the GPUs can fill up with CTAs waiting for each other's
 * producers -- a distributed deadlock that only the 2 s timeout breaks (seen at 32 lanes on 2 GPUs). */
int gpuhash_route_scatter_pub(const void *in_d, size_t n, int elem_words, uint32_t hash_mask_total, int log2_shards,
		const void *const *dst_ptrs, uint32_t *counts2_d, uint32_t *perm_d, size_t cap, int my_rank,
		const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs, uint32_t *ticket_d, uint32_t seq,
		const uint32_t *ack_flags_d, uint32_t *err_d, void *stream);
/* Tile-sorted routing (the default of megakv_b200/sharded.py): every CTA sorts its tile of 1024 requests by owner in shared
 * memory and writes each owner's run contiguously; map_d (gpuhash_route_map_bytes(cap) bytes, NULL for insert/delete
 * batches) records per request its position in the sorted tile and per tile the run starts, which is all
 * gpuhash_route_gather_tiles needs to bring the results (read run by run from the staging regions) back into request
 * order with coalesced stores.  Publication as in gpuhash_route_scatter_pub; waits are the caller's (gpuhash_wait_flags). */
size_t gpuhash_route_map_bytes(size_t cap);
int gpuhash_route_scatter_tiles(const void *in_d, size_t n, int elem_words, uint32_t hash_mask_total, int log2_shards,
		const void *const *dst_ptrs, uint32_t *counts2_d, void *map_d, size_t cap, int my_rank,
		const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs, uint32_t *ticket_d, uint32_t seq, void *stream);
int gpuhash_route_gather_tiles(const void *const *staged_ptrs, const void *map_d, size_t cap, int log2_shards,
		void *out_d, size_t n, void *stream);
int gpuhash_serve(const gpuhash_geom_t *g, void *table_d, int op, int log2_shards, const void *const *seg_in_ptrs,
		const uint32_t *seg_count_d, const void *const *seg_out_ptrs, size_t max_total, const uint32_t *req_flags_d, uint32_t *err_d,
		int my_rank, const void *const *peer_res_flag_ptrs, uint32_t *ticket_d, uint32_t seq, gpuhash_stats_t *stats_d, void *stream);
/* fused path: block the stream until flags_d[0..num) >= want (peers raise them with release semantics).  Done with stream
 * memory operations (no SM, graph-capturable) when the driver offers them (gpuhash_wait_mode() == 1), else by a one-CTA
 * kernel with a 2 s timeout (-> *err_d = 1).  GPUHASH_WAIT_MODE=kernel forces the latter. */
int gpuhash_wait_flags(const uint32_t *flags_d, int num, uint32_t want, uint32_t *err_d, void *stream);
int gpuhash_wait_mode(void);
int   gpuhash_ipc_export(void *dev_ptr, void *handle_out_64B);      /* cudaIpcGetMemHandle */
void *gpuhash_ipc_import(const void *handle_64B);                   /* cudaIpcOpenMemHandle, NULL on failure */
int   gpuhash_ipc_close(void *imported_ptr);

/* ---- sharded index, ONE kernel per scheduler cycle (megakv_b200/csrc/gpuhash_xchg.cu) ----
 * The routed path above as a software pipeline: launch j of a rank scatters its exchange j into the owners' inboxes,
 * serves exchange j-1 (what the peers sent one launch ago: search -> delete -> insert per source, results straight into
 * the origins' staging areas) and gathers exchange j-2 into its caller's search_out -- in ONE kernel whose warps take
 * interleaved scatter / lookup / gather tiles by ticket, so the routing traffic hides under the lookups' line fills.
 * Between GPUs there is one stream memory operation per launch (every peer's previous launch has raised its flag) and
 * no wait inside any kernel.  One exchange = one scheduler cycle of the reference (all its workers' batches,
 * src/mega_scheduler.c:393-504).  An xchg object owns a triple-buffered arena (inboxes, staging, routing maps, flags);
 * the arenas of all ranks are made known to each other with gpuhash_xchg_set_peers (plain device pointers inside one
 * process, CUDA IPC imports across processes: gpuhash_ipc_export / gpuhash_ipc_import on gpuhash_xchg_arena()).
 *   gpuhash_xchg_step   COLLECTIVE in the number of calls: every rank calls it equally often (empty parts allowed).
 *                       The buffers of call j may be device or pinned host memory; search_in/delete_in/insert_in must stay
 *                       valid until launch j has run, search_out receives the results when launch j+2 has run.
 *   gpuhash_xchg_flush  two steps without new requests: afterwards (and after synchronising the stream) every earlier
 *                       exchange is complete and its results are in place.
 *   gpuhash_xchg_error  0, a CUDA error, or -3 if a wait inside a kernel timed out.
 * Sequence numbers are baked into the launches: a CUDA graph that captured steps may be replayed once.
 * The kernel is persistent -- one CTA of 640 / 768 threads per SM, the whole register file -- and its delete / insert
 * tiles wait (bounded: 2 s, then gpuhash_xchg_error reports -3) until every lookup warp of the grid has run out of lookup
 * tiles, so all its CTAs must be able to become resident: one exchange object per GPU at a time, on a GPU the process has
 * to itself (no MPS partition smaller than the device).  Between ranks nothing inside a kernel ever waits. */
typedef struct gpuhash_xchg_s gpuhash_xchg_t;
gpuhash_xchg_t *gpuhash_xchg_create(const gpuhash_geom_t *g, void *table_d, uint32_t hash_mask_total, int log2_shards,
		int my_rank, size_t cap_search, size_t cap_update);
void    *gpuhash_xchg_arena(gpuhash_xchg_t *x, size_t *bytes);
int      gpuhash_xchg_set_peers(gpuhash_xchg_t *x, const void *const *peer_arenas);
int      gpuhash_xchg_set_stats(gpuhash_xchg_t *x, gpuhash_stats_t *stats_d);
int      gpuhash_xchg_step(gpuhash_xchg_t *x, const void *search_in, size_t n_search, void *search_out,
		const void *delete_in, size_t n_delete, const void *insert_in, size_t n_insert, void *stream);
int      gpuhash_xchg_flush(gpuhash_xchg_t *x, void *stream);
unsigned gpuhash_xchg_seq(const gpuhash_xchg_t *x);
int      gpuhash_xchg_error(gpuhash_xchg_t *x);
void     gpuhash_xchg_destroy(gpuhash_xchg_t *x);

/* ---- synthetic request streams generated on the device (bench tooling; SURVEY.md 8(d) key stream) ----
 * inserts: keys first..first+n-1 of the splitmix64 stream `seed`, loc = key index + 1; either output may be NULL.
 * queries: n searches for keys drawn from the first `population` keys, uniformly (theta 0) or Zipf(theta)
 * with zetan = sum_{k=1..population} k^-theta supplied by the caller; expect_loc_d (optional) gets index + 1. */
int gpuhash_gen_inserts(void *ielem_d, void *selem_d, uint64_t seed, uint64_t first, size_t n, void *stream);
int gpuhash_gen_queries(void *selem_d, void *expect_loc_d, uint64_t seed, uint64_t population, size_t n,
		uint64_t rng_seed, double theta, double zetan, void *stream);
int gpuhash_gen_requests(void *ielem_d, uint64_t seed, uint64_t population, size_t n,
		uint64_t rng_seed, double theta, double zetan, void *stream);   /* same draw as (sig, hash, loc) triples */

/* the same with ranks drawn by the REFERENCE's generator, bit for bit (src/zipf.h:44-183: mehcached_zipf_next with its
 * approximate pow and 48-bit LCG; what SURVEY.md 8(d) names for the zipf configs): elements first .. first+n-1 of the
 * sequence seeded with rand_seed (< 2^48); theta in [0, 1); zetan = mehcached_zeta(population, theta).  triples != 0:
 * (sig, hash, loc = rank + 1) records (insert / delete requests), else (sig, hash) searches. */
int gpuhash_gen_requests_ref_zipf(void *out_d, void *expect_loc_d, uint64_t seed, uint64_t population, size_t n,
		uint64_t rand_seed, uint64_t first, double theta, double zetan, int triples, void *stream);

/* ---- timed loops (CUDA events on the launching streams; the Python bench only orchestrates) ---- */
typedef struct gpuhash_bench_result_s {
	float  total_ms;        /* first launch -> last completion, events */
	float  search_ms;       /* summed per-launch device time of the search kernels (serial leg only) */
	unsigned long long launches;
	unsigned long long search_ops, insert_ops, delete_ops;
	unsigned long long h2d_bytes, d2h_bytes;
} gpuhash_bench_result_t;

/* Device-resident mixed workload: `steps` batches laid out back to back in device memory
 * (search_d: steps*n_search selem_t, insert_d: steps*n_insert ielem_t, out_d: steps*2*n_search loc_t),
 * issued round-robin over `streams` streams, one search launch + one insert launch per batch.
 * use_graph != 0: the whole issue sequence is captured once into a CUDA graph and the graph launch is
 * what gets timed (removes the host's per-launch cost, not the device's). */
int gpuhash_bench_resident(const gpuhash_geom_t *g, void *table_d,
		const void *search_d, size_t n_search, void *out_d,
		const void *insert_d, size_t n_insert,
		int steps, int streams, int use_graph, gpuhash_bench_result_t *res);

/* Same workload through gpuhash_index_submit with pinned host buffers (H2D + D2H inside the timing). */
int gpuhash_bench_e2e(gpuhash_index_t *ix,
		const void *search_h, size_t n_search, void *out_h,
		const void *insert_h, size_t n_insert,
		int steps, int use_graph, gpuhash_bench_result_t *res);

/* Whole scheduler cycles, ONE launch per step (gpuhash_cycle_multi_ex) over `batches` worker batches: resident in device
 * memory and timed by CUDA events (bench_cycles; steps round-robin over 1..4 streams), or end to end from pinned host
 * memory through gpuhash_index_submit_all / gpuhash_index_wait with at most `depth` cycles in flight, timed by the HOST'S
 * WALL CLOCK from the first submit to the return of the last wait (bench_e2e_cycles).  Batch b of step i is batch
 * i * batches + b of the arrays (e2e: modulo host_batches). */
int gpuhash_bench_cycles(const gpuhash_geom_t *g, void *table_d,
		const void *search_d, size_t n_search, void *out_d,
		const void *insert_d, size_t n_insert,
		int batches, int steps, int streams, gpuhash_bench_result_t *res);
int gpuhash_bench_e2e_cycles(gpuhash_index_t *ix,
		const void *search_h, size_t n_search, void *out_h,
		const void *insert_h, size_t n_insert,
		int batches, size_t host_batches, int steps, int depth, gpuhash_bench_result_t *res);

/* The same cycles through gpuhash_ring_submit (no launches: total_ms is host wall clock, first doorbell -> last
 * completion mark); rtt_us (optional) = median round trip of rtt_reps isolated search batches. */
int gpuhash_bench_ring(gpuhash_ring_t *q, const void *search_h, size_t n_search, void *out_h,
		const void *insert_h, size_t n_insert, int steps, gpuhash_bench_result_t *res, int rtt_reps, float *rtt_us);

#ifdef __cplusplus
}
#endif
#endif