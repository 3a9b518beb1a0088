/*
 * gpuhash_oracle.h -- single-threaded CPU restatement of Mega-KV's GPU hash index.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py may load this library.
 * Nothing under megakv_b200/ links, imports or calls it; the product path has
 * no CPU fallback.
 *
 * What it restates (reference = pzrq/megakv, paths relative to its root):
 *   libgpuhash/gpu_hash.h:38-104    types, HASH_MASK, BLOCK_HASH_MASK, policies
 *   libgpuhash/gpu_hash.cu:28-75    hash_search
 *   libgpuhash/gpu_hash.cu:77-229   hash_insert_2choice
 *   libgpuhash/gpu_hash.cu:231-433  hash_insert_cuckoo
 *   libgpuhash/gpu_hash.cu:435-480  hash_delete
 * applied one request at a time in batch order (insert segments in segment
 * order), i.e. the result the reference kernels produce when no two requests
 * of a launch race on a bucket.
 *
 * Pin status: the reference ships no golden vectors for this path (its only
 * check is the inserted=>findable / deleted=>not-findable property of
 * libgpuhash/test/insert_test.c:178-195,237-244, unseeded).  This oracle is
 * pinned by (1) that property, (2) the regression anchors of SURVEY.md
 * Appendix D, (3) the fixture shape of libgpuhash/test/back/py_search_stream.c:
 * 104-129, and (4) -- when oracle/_ref has been built and a GPU is present --
 * a differential run against the reference's own kernels compiled from
 * /root/reference in legacy-warp mode (tests/test_ref_differential.py, golden
 * vectors under tests/golden/ref_*.npz).  See DESIGN.md "Oracle pin".
 */
#ifndef GPUHASH_ORACLE_H
#define GPUHASH_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_CUCKOO   0
#define ORC_2CHOICE  1

typedef struct orc_geom_s {
	uint32_t hash_mask;    /* HASH_MASK        gpu_hash.h:61 */
	uint32_t block_mask;   /* BLOCK_HASH_MASK  gpu_hash.h:69 */
	int      algo;         /* ORC_CUCKOO | ORC_2CHOICE  (gpu_hash.h:72-73) */
	int      max_cuckoo;   /* MAX_CUCKOO_NUM   gpu_hash.h:75 */
} orc_geom_t;

typedef struct orc_sel_s { uint32_t sig, hash; } orc_sel_t;          /* selem_t */
typedef struct orc_iel_s { uint32_t sig, hash, loc; } orc_iel_t;     /* ielem_t / delem_t */

/* what happened to a batch of inserts (the reference reports nothing) */
typedef struct orc_stats_s {
	uint64_t skipped;       /* sig==0 && loc==0            gpu_hash.cu:101,259 */
	uint64_t updated;       /* signature already present -> loc overwritten */
	uint64_t placed_b1;     /* claimed an empty slot of bucket 1 */
	uint64_t placed_b2;     /* claimed an empty slot of an alternate bucket */
	uint64_t to_b2;         /* requests that left bucket 1 */
	uint64_t displaced;     /* victims captured and re-homed (cuckoo, c < max) */
	uint64_t dropped;       /* victims overwritten at c == max (cuckoo) */
	uint64_t overwritten;   /* 2-choice: sig-only overwrite of a full bucket 2 */
	uint64_t chain_hist[8]; /* cuckoo: requests by number of displacements 0..6 */
} orc_stats_t;

void     orc_geom_init(orc_geom_t *g, int mem_p, int algo);
size_t   orc_table_bytes(const orc_geom_t *g);
uint32_t orc_bucket1(const orc_geom_t *g, uint32_t hash);
uint32_t orc_bucket2(const orc_geom_t *g, uint32_t hash, uint32_t sig);

/* out[2n] must be pre-zeroed by the caller: only hits are stored. */
void     orc_search(const void *table, const orc_geom_t *g,
		const orc_sel_t *in, size_t n, uint32_t *out);
/* orc_search with the buckets of a later request prefetched (same results; the CPU baseline's loop) */
void     orc_search_pf(const void *table, const orc_geom_t *g,
		const orc_sel_t *in, size_t n, uint32_t *out);
void     orc_insert(void *table, const orc_geom_t *g,
		const orc_iel_t *in, size_t n, orc_stats_t *st /* may be NULL, accumulates */);
void     orc_insert_blocks(void *table, const orc_geom_t *g,
		const orc_iel_t *const *blk, const int *blk_n, int num_blks, orc_stats_t *st);
/* returns the number of slots whose signature was zeroed */
uint64_t orc_delete(void *table, const orc_geom_t *g, const orc_iel_t *in, size_t n);

/* table inspection: occupied slot count and an order-independent digest of the
 * multiset {(sig,loc) : sig != 0}.  digest[0] = sum, digest[1] = xor of a
 * 64-bit mix of each pair; per_bucket != 0 mixes the bucket index in as well. */
uint64_t orc_table_occupied(const void *table, const orc_geom_t *g);
void     orc_table_digest(const void *table, const orc_geom_t *g, int per_bucket,
		uint64_t digest[2]);

/* key stream of SURVEY.md 8(d): key_i = i-th output (i from 0) of splitmix64
 * started at state `seed`; sig = low 32 bits (0 -> 1), hash = high 32 bits,
 * loc = first_index + i + 1. */
/* the reference's key-rank generator (src/zipf.h:26-183, Gray et al. with an approximate pow and a 48-bit LCG),
 * restated; pinned to vectors recorded from zipf.h itself (tests/golden/zipf_ref.npz, tests/golden/make_zipf_golden.py) */
typedef struct orc_zipf_s {
	uint64_t n; double theta, alpha, thres, dbl_n, zetan, eta; uint64_t rand_state;
} orc_zipf_t;
void     orc_zipf_init(orc_zipf_t *z, uint64_t n, double theta, uint64_t rand_seed);
void     orc_zipf_init_zetan(orc_zipf_t *z, uint64_t n, double theta, uint64_t rand_seed, double zetan);
uint64_t orc_zipf_next(orc_zipf_t *z);
void     orc_zipf_fill(orc_zipf_t *z, size_t n, uint64_t *ranks);
uint64_t orc_splitmix64(uint64_t *state);
void     orc_keys_fill(uint64_t seed, uint64_t first_index, size_t n,
		orc_iel_t *iel /* may be NULL */, orc_sel_t *sel /* may be NULL */);

/* multi-threaded drivers for the CPU baseline.  Inserts/deletes are partitioned
 * by the top IBLOCK_P bits of the bucket index (closed under the alternate
 * bucket function, so partitions never touch each other: at most 8 threads);
 * searches are split by index range (any thread count). */
void     orc_search_mt(const void *table, const orc_geom_t *g,
		const orc_sel_t *in, size_t n, uint32_t *out, int threads);
void     orc_insert_mt(void *table, const orc_geom_t *g,
		const orc_iel_t *in, size_t n, int threads);
double   orc_now_sec(void);

/* Persistent thread pool for the CPU arm of bench.py: threads live across cycles and are handed work through barriers.
 * orc_pool_cycle = one scheduler cycle of `batches` worker batches (mega_scheduler.c:393-504) on the host: all searches
 * (index ranges over all threads, `out` zeroed inside like the caller's memset at :406), then all inserts (the 8 closed
 * bucket ranges over min(threads, 8) threads, batch order kept).  The table ends equal to orc_insert() over the batches. */
typedef struct orc_pool_s orc_pool_t;
orc_pool_t *orc_pool_create(int threads);
int      orc_pool_threads(const orc_pool_t *p);
void     orc_pool_destroy(orc_pool_t *p);
void     orc_pool_cycle(orc_pool_t *p, void *table, const orc_geom_t *g, const orc_sel_t *sel, size_t n_search,
		uint32_t *out, const orc_iel_t *iel, size_t n_insert, int batches);
void     orc_pool_insert(orc_pool_t *p, void *table, const orc_geom_t *g, const orc_iel_t *iel, size_t n);
void     orc_pool_preload(orc_pool_t *p, void *table, const orc_geom_t *g, uint64_t seed, uint64_t first, uint64_t count);
void     orc_gen_queries(uint64_t seed, uint64_t population, size_t n, uint64_t rng_seed, orc_sel_t *out);

#ifdef __cplusplus
}
#endif
#endif
