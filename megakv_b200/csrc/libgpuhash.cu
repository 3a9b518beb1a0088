/*
 * libgpuhash.cu -- host side of the launches: the legacy C ABI (libgpuhash.h), the
 * run-time-geometry *_ex calls (gpuhash_ex.h), the device/stream plumbing and the
 * random-sector roofline probe.
 *
 * Built without exceptions/RTTI and without any libstdc++ symbol, so the static
 * archive links with plain `gcc ... -lgpuhash -lcudart` exactly like the reference's
 * (src/Makefile:26; src/mega.c:23-27 fakes __gxx_personality_v0 for the same reason).
 *
 * There is no CPU fallback anywhere in this file: every operation is a kernel launch.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "libgpuhash.h"
#include "gpuhash_ex.h"
#define GH_DEFINE_KERNELS
#include "gpuhash_kernels.cuh"

static_assert(sizeof(bucket_t) == 64 && sizeof(gh::Bucket) == 64, "bucket_t must be 64 B (gpu_hash.h:79-82)");
static_assert(sizeof(selem_t) == 8 && sizeof(ielem_t) == 12, "request layouts (gpu_hash.h:85-104)");
static_assert(sizeof(gpuhash_geom_t) == sizeof(gh::Geom), "geom mirror");
static_assert(sizeof(gpuhash_stats_t) == sizeof(gh::Stats), "stats mirror");

#if defined(HASH_2CHOICE)
#  define GH_DEFAULT_ALGO GPUHASH_2CHOICE
#else
#  define GH_DEFAULT_ALGO GPUHASH_CUCKOO
#endif

/* process-wide defaults of the legacy entry points = the header this file was compiled with */
#if defined(GPUHASH_DEFAULT_LAYOUT_REFERENCE)
#  define GH_DEFAULT_LAYOUT GPUHASH_LAYOUT_REFERENCE
#else
#  define GH_DEFAULT_LAYOUT GPUHASH_LAYOUT_PAIRS
#endif
static gpuhash_geom_t g_default_geom = {
	(uint32_t)HASH_MASK, (uint32_t)BLOCK_HASH_MASK, GH_DEFAULT_ALGO, 5u, GH_DEFAULT_LAYOUT
};
static gpuhash_tune_t g_tune = { 0, 0, 4, 1 };

static int g_sm_count[64];          /* 0 = not queried yet */
static int g_l2_bytes[64];

static int sm_count_now(void)
{
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
	if (g_sm_count[dev] == 0) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
		g_sm_count[dev] = n;
	}
	return g_sm_count[dev];
}

static size_t l2_bytes_now(void)
{
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (size_t)126 << 20;
	if (g_l2_bytes[dev] == 0) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrL2CacheSize, dev) != cudaSuccess || n <= 0) n = 126 << 20;
		g_l2_bytes[dev] = n;
	}
	return (size_t)g_l2_bytes[dev];
}

/* GPUHASH_SEARCH_STAGED=0 in the environment keeps the default search on the plain four-lane kernel (A/B runs) */
static int staged_default(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("GPUHASH_SEARCH_STAGED"); v = (e && e[0] == '0') ? 0 : 1; }
	return v;
}
/* GPUHASH_UPDATE_PAIR=0: insert/delete launches keep one thread per request (A/B runs; default: two lanes per request
 * for the pair layout, one L2 request per bucket) */
static int update_pair_default(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("GPUHASH_UPDATE_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
	return v;
}
/* GPUHASH_PDL=1: the search and insert launches of a stream are chained by programmatic dependent launch -- the next
 * kernel's CTAs are scheduled while the previous kernel still runs and wait (griddepcontrol.wait) for its completion, so
 * the launch latency between the small kernels of a 64 K batch is hidden.  Captured into CUDA graphs as programmatic edges. */
static int pdl_on(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("GPUHASH_PDL"); v = (e && e[0] == '1') ? 1 : 0; }
	return v;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned blocks, unsigned threads, cudaStream_t s, Args... args)
{
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	at[0].val.programmaticStreamSerializationAllowed = pdl_on();
	cfg.attrs = at; cfg.numAttrs = 1;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
/* GPUHASH_SEARCH_QPT=<n>: process-wide override of tune.search_qpt == 0 (A/B runs of the launch shapes) */
static int qpt_env(void)
{
	static int v = 1 << 30;
	if (v == 1 << 30) { const char *e = getenv("GPUHASH_SEARCH_QPT"); v = e && *e ? atoi(e) : 0; }
	return v;
}

static inline gh::Geom to_geom(const gpuhash_geom_t *g)
{
	gh::Geom r; r.hash_mask = g->hash_mask; r.block_mask = g->block_mask; r.algo = g->algo; r.max_cuckoo = g->max_cuckoo;
	r.layout = g->layout;
	return r;
}

/* ------------------------------------------------------------------ geometry */

extern "C" int gpuhash_geom_init(gpuhash_geom_t *g, int mem_p, unsigned algo)
{
	return gpuhash_geom_init_shard(g, mem_p, 0, algo);
}

/* gpuhash_geom_init with the layout chosen for the table: {sig, loc} pairs (every commit one 64-bit CAS; the fewest L2
 * requests per probe beyond L2) unless the table is L2-resident AND the policy is two-choice -- then the reference's own
 * byte layout, whose one-thread search reads the signature rows and, on a hit only, the location word: fewest L2 sectors,
 * which is what an L2-resident table pays for (76 vs 67 Gops/s at MEM_P 26).  Cuckoo tables keep the pair
 * layout at any size: only there an eviction chain re-homes a (sig, loc) pair atomically under concurrent inserts. */
extern "C" int gpuhash_geom_init_auto(gpuhash_geom_t *g, int mem_p, unsigned algo)
{
	int rc = gpuhash_geom_init_shard(g, mem_p, 0, algo);
	if (rc) return rc;
	if (algo == GPUHASH_2CHOICE && gpuhash_table_bytes(g) <= l2_bytes_now()) g->layout = GPUHASH_LAYOUT_REFERENCE;
	return 0;
}

extern "C" int gpuhash_geom_init_shard(gpuhash_geom_t *g, int mem_p_total, int log2_shards, unsigned algo)
{
	/* BUC_P = 6, IBLOCK_P = 3 (gpu_hash.h:57,67).  hash_t is 32 bits => at most 2^32 buckets, MEM_P <= 38;
	 * the alternate bucket keeps the top 3 bits of the bucket index, so at most 8 shards stay closed. */
	if (!g || algo > GPUHASH_2CHOICE || log2_shards < 0 || log2_shards > 3) return -1;
	if (mem_p_total < 6 + 3 || mem_p_total > 38) return -1;
	g->hash_mask  = (uint32_t)((1ULL << (mem_p_total - 6 - log2_shards)) - 1);
	g->block_mask = (uint32_t)((1ULL << (mem_p_total - 6 - 3)) - 1);
	g->algo = algo;
	g->max_cuckoo = 5;
	g->layout = GPUHASH_LAYOUT_PAIRS;
	return 0;
}

extern "C" size_t gpuhash_table_bytes(const gpuhash_geom_t *g)
{
	return ((size_t)g->hash_mask + 1) * sizeof(gh::Bucket);
}

extern "C" void gpuhash_set_default_geom(const gpuhash_geom_t *g) { g_default_geom = *g; }
extern "C" void gpuhash_get_default_geom(gpuhash_geom_t *g) { *g = g_default_geom; }
extern "C" void gpuhash_set_tuning(const gpuhash_tune_t *t) { g_tune = *t; }
extern "C" void gpuhash_get_tuning(gpuhash_tune_t *t) { *t = g_tune; }

/* ------------------------------------------------------------------ launches */

template <int kQpt, int kMode>
static void launch_search_mode(const uint2 *in, uint2 *out, const gh::Bucket *table, size_t n,
		const gh::Geom &g, gh::Stats *st, cudaStream_t s)
{
	size_t blocks = (n + (size_t)256 * kQpt - 1) / ((size_t)256 * kQpt);
	if (blocks > 0x7fffffffULL) blocks = 0x7fffffffULL;
	gh::search_kernel<kQpt, kMode><<<(unsigned)blocks, 256, 0, s>>>(in, out, table, n, g, st);
}

template <int kQpt>
static void launch_search(int mode, const uint2 *in, uint2 *out, const gh::Bucket *table, size_t n,
		const gh::Geom &g, gh::Stats *st, cudaStream_t s)
{
	if (mode == gh::kSearchPairs)           launch_search_mode<kQpt, gh::kSearchPairs>(in, out, table, n, g, st, s);
	else if (mode == gh::kSearchSplitWhole) launch_search_mode<kQpt, gh::kSearchSplitWhole>(in, out, table, n, g, st, s);
	else                                    launch_search_mode<kQpt, gh::kSearchSplitLazy>(in, out, table, n, g, st, s);
}

extern "C" int gpuhash_search_ex(const gpuhash_geom_t *g, const void *selem_d, void *out_d,
		const void *table_d, size_t n, gpuhash_stats_t *stats_d, void *stream)
{
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || (n && (!selem_d || !out_d || !table_d))) return -1;
	if (n == 0) return 0;
	int qpt = g_tune.search_qpt;
	if (qpt == 0) qpt = qpt_env();
	if (qpt == 0) {
		/* Beyond L2 the scarce resource is L2 requests for non-resident 128 B lines (~47 G/s on B200), and the
		 * two sectors of a bucket cost ONE request only when two lanes ask for them in the same instruction
		 * (profiles/r01_l2_requests.md): four lanes per request.  The pair layout needs whole buckets anyway,
		 * so it always takes that shape.  The reference layout on an L2-resident table is cheapest with one
		 * thread per request reading the signature rows and, on a hit only, the location word. */
		if (g->layout == GPUHASH_LAYOUT_PAIRS || gpuhash_table_bytes(g) > l2_bytes_now()) qpt = staged_default() ? -6 : -4;
		else qpt = n <= (size_t)sm_count_now() * 2048 ? 1 : 2;
	}
	const uint2 *in = (const uint2 *)selem_d; uint2 *out = (uint2 *)out_d;
	const gh::Bucket *t = (const gh::Bucket *)table_d; gh::Stats *st = (gh::Stats *)stats_d;
	cudaStream_t s = (cudaStream_t)stream;
	gh::Geom gg = to_geom(g);
	if ((qpt == -5 || qpt == -6) && g_tune.search_qpt == 0 && qpt_env() == 0 && n >= ((size_t)1 << 20)) {
		/* Chosen by default, not asked for: one launch over a very large DEVICE-resident batch runs ~10 % faster with the
		 * plain four-lane kernel and 64 CTAs per SM (19.9 vs 17.7-17.9 Gops/s at 2^24 requests); batches in pinned host
		 * memory always go tile-wise (512 B transactions over the host link). */
		cudaPointerAttributes pa;
		if (cudaPointerGetAttributes(&pa, in) == cudaSuccess && pa.type == cudaMemoryTypeDevice &&
		    cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeDevice) qpt = -4;
		(void)cudaGetLastError();
	}
	if (qpt == -6 && ((uintptr_t)in & 7u) == 0 && ((uintptr_t)out & 7u) == 0) {
		/* four lanes per request, every warp on its own: 512 B tile in, four loads per lane in flight, 512 B tile out */
		const unsigned head = ((uintptr_t)in & 15u) ? 1u : 0u;
		const int out_vec = (((uintptr_t)out + 8u * head) & 15u) == 0;
		size_t tiles = (n - head + gh::kTileReq - 1) / gh::kTileReq;
		/* a 64 K batch is 973 tiles: with 8 warps per CTA that is 122 CTAs -- fewer than SMs.  Small launches take
		 * smaller CTAs so that a lone batch still reaches every SM; large ones 8 warps per CTA, 8 CTAs per SM. */
		static int blk_env = -1;
		if (blk_env < 0) { const char *e = getenv("GPUHASH_WARP_BLOCK"); blk_env = e && *e ? atoi(e) : 0; }
		unsigned threads = blk_env ? (unsigned)blk_env : (tiles <= (size_t)sm_count_now() * 16 ? 64u : 256u);
		const size_t wpc = threads / 32;
		size_t blocks = (tiles + wpc - 1) / wpc, cap = (size_t)sm_count_now() * (2048 / threads);
		if (blocks > cap) blocks = cap;
		if (blocks == 0) blocks = 1;
		if (g->layout == GPUHASH_LAYOUT_PAIRS) launch_pdl(gh::search_warp_kernel<true, false>, (unsigned)blocks, threads, s, in, (void *)out, t, n, gg, st, head, out_vec);
		else                                   launch_pdl(gh::search_warp_kernel<false, false>, (unsigned)blocks, threads, s, in, (void *)out, t, n, gg, st, head, out_vec);
		return (int)cudaGetLastError();
	}
	if (qpt == -5 && ((uintptr_t)in & 7u) == 0 && ((uintptr_t)out & 7u) == 0) {
		/* four lanes per request, batch staged through shared memory by 512 B bulk copies (zero-copy host buffers,
		 * fewer L2 requests for the streams).  Tiles start at the first 16 B-aligned request. */
		const unsigned head = ((uintptr_t)in & 15u) ? 1u : 0u;
		const int out_bulk = (((uintptr_t)out + 8u * head) & 15u) == 0;
		size_t tiles = (n - head + gh::kTileReq - 1) / gh::kTileReq;
		size_t cap = (size_t)sm_count_now() * 8;
		if (tiles > cap) tiles = cap;
		if (tiles == 0) tiles = 1;
		if (g->layout == GPUHASH_LAYOUT_PAIRS) gh::search_quad_staged_kernel<true><<<(unsigned)tiles, 256, 0, s>>>(in, out, t, n, gg, st, head, out_bulk);
		else                                   gh::search_quad_staged_kernel<false><<<(unsigned)tiles, 256, 0, s>>>(in, out, t, n, gg, st, head, out_bulk);
		return (int)cudaGetLastError();
	}
	if (qpt == -4 || qpt == -5 || qpt == -6) {       /* four lanes per request, one L2 request per bucket */
		size_t blocks = (n * 4 + 255) / 256;
		size_t cap = (size_t)sm_count_now() * 64;
		if (blocks > cap) blocks = cap;
		if (g->layout == GPUHASH_LAYOUT_PAIRS) gh::search_quad_kernel<true><<<(unsigned)blocks, 256, 0, s>>>(in, out, t, n, gg, st);
		else                                   gh::search_quad_kernel<false><<<(unsigned)blocks, 256, 0, s>>>(in, out, t, n, gg, st);
		return (int)cudaGetLastError();
	}
	int mode = gh::kSearchPairs;
	if (g->layout == GPUHASH_LAYOUT_REFERENCE) {
		/* in HBM a probe costs a whole 128 B line whatever is asked of it, so take the location row with the
		 * signature row; in L2 sectors are the cost, so fetch the location word only on a hit */
		if (g_tune.search_split_mode == 1) mode = gh::kSearchSplitLazy;
		else if (g_tune.search_split_mode == 2) mode = gh::kSearchSplitWhole;
		else mode = gh::kSearchSplitLazy;
		if (qpt == -1) {                             /* comparison shape: 4 lanes per request, 128-bit loads */
			size_t blocks = (n * 4 + 255) / 256;
			size_t cap = (size_t)sm_count_now() * 64;
			if (blocks > cap) blocks = cap;
			gh::search_coop4_kernel<<<(unsigned)blocks, 256, 0, s>>>(in, out, t, n, gg);
			return (int)cudaGetLastError();
		}
	}
	if (qpt >= 4)      launch_search<4>(mode, in, out, t, n, gg, st, s);
	else if (qpt >= 2) launch_search<2>(mode, in, out, t, n, gg, st, s);
	else               launch_search<1>(mode, in, out, t, n, gg, st, s);
	return (int)cudaGetLastError();
}

/* One word per request: out32[i] = bucket-1 hit, else bucket-2 hit, else 0 -- what the sender picks out of the pair
 * (src/mega_send.c:411-414), decided on the device so that half the result bytes cross the host link. */
extern "C" int gpuhash_search_compact_ex(const gpuhash_geom_t *g, const void *selem_d, void *out32_d,
		const void *table_d, size_t n, gpuhash_stats_t *stats_d, void *stream)
{
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || (n && (!selem_d || !out32_d || !table_d))) return -1;
	if (((uintptr_t)selem_d & 7u) || ((uintptr_t)out32_d & 3u)) return -1;
	if (n == 0) return 0;
	const uint2 *in = (const uint2 *)selem_d;
	const unsigned head = ((uintptr_t)in & 15u) ? 1u : 0u;
	const int out_vec = (((uintptr_t)out32_d + 4u * head) & 7u) == 0;
	size_t tiles = (n - head + gh::kTileReq - 1) / gh::kTileReq;
	size_t blocks = (tiles + 7) / 8, cap = (size_t)sm_count_now() * 8;
	if (blocks > cap) blocks = cap;
	if (blocks == 0) blocks = 1;
	gh::Geom gg = to_geom(g);
	cudaStream_t s = (cudaStream_t)stream;
	if (g->layout == GPUHASH_LAYOUT_PAIRS)
		gh::search_warp_kernel<true, true><<<(unsigned)blocks, 256, 0, s>>>(in, out32_d, (const gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d, head, out_vec);
	else
		gh::search_warp_kernel<false, true><<<(unsigned)blocks, 256, 0, s>>>(in, out32_d, (const gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d, head, out_vec);
	return (int)cudaGetLastError();
}

/* Key bytes -> search requests (src/mega_recv.c:349-362): n keys of nkey >= 8 bytes, `stride` bytes apart; fold != 0 =
 * the reference built with -DSIGNATURE (XOR of all 8-byte words, tail masked), fold == 0 = first 8 bytes only. */
extern "C" int gpuhash_fold_keys_ex(const void *keys_d, size_t stride, unsigned nkey, int fold, size_t n, void *selem_out_d, void *stream)
{
	if (nkey < 8 || stride < nkey || (n && (!keys_d || !selem_out_d)) || ((uintptr_t)selem_out_d & 7u)) return -1;
	if (n == 0) return 0;
	size_t blocks = (n + 255) / 256, cap = (size_t)sm_count_now() * 16;
	if (blocks > cap) blocks = cap;
	gh::fold_keys_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char *)keys_d, stride, nkey, fold, n, (uint2 *)selem_out_d);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_insert_ex(const gpuhash_geom_t *g, void *table_d, const void *const *blk_input_d,
		const int *blk_elem_num_d, int num_blks, gpuhash_stats_t *stats_d, unsigned flags, void *stream)
{
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || num_blks < 0 || (num_blks && (!table_d || !blk_input_d || !blk_elem_num_d))) return -1;
	if (num_blks == 0) return 0;
	cudaStream_t s = (cudaStream_t)stream;
	gh::Geom gg = to_geom(g);
	if (flags & GPUHASH_INSERT_SERIAL) {
		if (gg.layout == gh::kLayoutPairs)
			gh::insert_serial_kernel<true><<<1, 32, 0, s>>>((gh::Bucket *)table_d, (const uint32_t *const *)blk_input_d,
					blk_elem_num_d, num_blks, nullptr, 0, gg, (gh::Stats *)stats_d);
		else
			gh::insert_serial_kernel<false><<<1, 32, 0, s>>>((gh::Bucket *)table_d, (const uint32_t *const *)blk_input_d,
					blk_elem_num_d, num_blks, nullptr, 0, gg, (gh::Stats *)stats_d);
	} else {
		int per_sm = g_tune.insert_ctas_per_sm > 0 ? g_tune.insert_ctas_per_sm : 4;
		unsigned blocks = (unsigned)(sm_count_now() * per_sm);
		if (gg.layout == gh::kLayoutPairs)
			gh::insert_segments_kernel<true><<<blocks, 256, 0, s>>>((gh::Bucket *)table_d,
					(const uint32_t *const *)blk_input_d, blk_elem_num_d, num_blks, gg, (gh::Stats *)stats_d);
		else
			gh::insert_segments_kernel<false><<<blocks, 256, 0, s>>>((gh::Bucket *)table_d,
					(const uint32_t *const *)blk_input_d, blk_elem_num_d, num_blks, gg, (gh::Stats *)stats_d);
	}
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_insert_flat_ex(const gpuhash_geom_t *g, void *table_d, const void *ielem_d, size_t n,
		gpuhash_stats_t *stats_d, unsigned flags, void *stream)
{
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || (n && (!table_d || !ielem_d))) return -1;
	if (n == 0) return 0;
	cudaStream_t s = (cudaStream_t)stream;
	gh::Geom gg = to_geom(g);
	if (flags & GPUHASH_INSERT_SERIAL) {
		if (gg.layout == gh::kLayoutPairs)
			gh::insert_serial_kernel<true><<<1, 32, 0, s>>>((gh::Bucket *)table_d, nullptr, nullptr, 0,
					(const uint32_t *)ielem_d, n, gg, (gh::Stats *)stats_d);
		else
			gh::insert_serial_kernel<false><<<1, 32, 0, s>>>((gh::Bucket *)table_d, nullptr, nullptr, 0,
					(const uint32_t *)ielem_d, n, gg, (gh::Stats *)stats_d);
	} else {
		size_t blocks = (n + 255) / 256;
		size_t cap = (size_t)sm_count_now() * 32;        /* grid-stride beyond 32 CTAs per SM */
		if (gg.layout == gh::kLayoutPairs && update_pair_default()) {
			blocks = (2 * n + 255) / 256;
			if (blocks > cap) blocks = cap;
			launch_pdl(gh::insert_flat_pair_kernel, (unsigned)blocks, 256u, s, (gh::Bucket *)table_d, (const uint32_t *)ielem_d, n, gg, (gh::Stats *)stats_d);
			return (int)cudaGetLastError();
		}
		if (blocks > cap) blocks = cap;
		if (gg.layout == gh::kLayoutPairs)
			gh::insert_flat_kernel<true><<<(unsigned)blocks, 256, 0, s>>>((gh::Bucket *)table_d,
					(const uint32_t *)ielem_d, n, gg, (gh::Stats *)stats_d);
		else
			gh::insert_flat_kernel<false><<<(unsigned)blocks, 256, 0, s>>>((gh::Bucket *)table_d,
					(const uint32_t *)ielem_d, n, gg, (gh::Stats *)stats_d);
	}
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_delete_ex(const gpuhash_geom_t *g, const void *delem_d, void *table_d, size_t n,
		gpuhash_stats_t *stats_d, unsigned flags, void *stream)
{
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || (n && (!table_d || !delem_d))) return -1;
	if (n == 0) return 0;
	cudaStream_t s = (cudaStream_t)stream;
	gh::Geom gg = to_geom(g);
	if (flags & GPUHASH_INSERT_SERIAL) {
		if (gg.layout == gh::kLayoutPairs)
			gh::delete_serial_kernel<true><<<1, 32, 0, s>>>((const uint32_t *)delem_d, (gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d);
		else
			gh::delete_serial_kernel<false><<<1, 32, 0, s>>>((const uint32_t *)delem_d, (gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d);
	} else {
		size_t blocks = (n + 255) / 256;
		size_t cap = (size_t)sm_count_now() * 32;
		if (gg.layout == gh::kLayoutPairs && update_pair_default()) {
			blocks = (2 * n + 255) / 256;
			if (blocks > cap) blocks = cap;
			gh::delete_pair_kernel<<<(unsigned)blocks, 256, 0, s>>>((const uint32_t *)delem_d, (gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d);
			return (int)cudaGetLastError();
		}
		if (blocks > cap) blocks = cap;
		if (gg.layout == gh::kLayoutPairs)
			gh::delete_kernel<true><<<(unsigned)blocks, 256, 0, s>>>((const uint32_t *)delem_d, (gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d);
		else
			gh::delete_kernel<false><<<(unsigned)blocks, 256, 0, s>>>((const uint32_t *)delem_d, (gh::Bucket *)table_d, n, gg, (gh::Stats *)stats_d);
	}
	return (int)cudaGetLastError();
}

/* Rewrites the table in place from its current layout (g->layout) to `to_layout`.  The caller updates g->layout. */
extern "C" int gpuhash_table_convert(const gpuhash_geom_t *g, void *table_d, unsigned to_layout, void *stream)
{
	if (!g || !table_d || g->layout > GPUHASH_LAYOUT_REFERENCE || to_layout > GPUHASH_LAYOUT_REFERENCE) return -1;
	if (g->layout == to_layout) return 0;
	size_t buckets = (size_t)g->hash_mask + 1;
	size_t blocks = (buckets + 255) / 256, cap = (size_t)sm_count_now() * 32;
	if (blocks > cap) blocks = cap;
	gh::convert_layout_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((gh::Bucket *)table_d, buckets,
			to_layout == GPUHASH_LAYOUT_PAIRS);
	return (int)cudaGetLastError();
}

/* ---- one launch per scheduler cycle (gh::cycle_multi_kernel) ----
 * Workspaces: the kernel orders its phases through a few counters in device memory that must be zero at launch and are
 * left zero at exit, so one workspace serves one launch at a time.  Callers that replay launches from CUDA graphs, or
 * issue from several host threads, own their workspaces (gpuhash_cycle_multi_ex; gpuhash_index_* keeps one per worker
 * stream and one per submit_all slot).  The stateless entry points (gpuhash_cycle_ex, the legacy gpu_delete_insert) take a
 * slot from a per-device pool by an atomic round-robin counter: a slot is reused GH_CYCLE_SLOTS launches later. */
#define GH_CYCLE_SLOTS 4096
#define GH_POOL_WS_WORDS 64                       /* gh::cycle_workspace_words(1) = 48, rounded to 256 B */
static unsigned int *g_cycle_pool[64];            /* per device: GH_CYCLE_SLOTS x GH_POOL_WS_WORDS words, zero */
static unsigned int g_cycle_next[64];
static volatile unsigned int *g_cycle_err[64];    /* per device: pinned word a cycle kernel sets when a phase wait timed out */

/* Per-device state of the library (the workspace pool, cached device attributes).  Idempotent; synchronises. */
extern "C" int gpuhash_init_device(void)
{
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
	(void)sm_count_now(); (void)l2_bytes_now();
	if (!__atomic_load_n(&g_cycle_pool[dev], __ATOMIC_ACQUIRE)) {
		unsigned int *p = NULL;
		const size_t bytes = (size_t)GH_CYCLE_SLOTS * GH_POOL_WS_WORDS * sizeof(unsigned int);
		cudaError_t e = cudaMalloc((void **)&p, bytes);
		if (e != cudaSuccess) return (int)e;
		/* cudaMemset on device memory is asynchronous to the host and the library's streams are non-blocking:
		 * nothing may be launched on the pool before the fill has finished */
		if ((e = cudaMemset(p, 0, bytes)) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { cudaFree(p); return (int)e; }
		unsigned int *err = NULL;
		if ((e = cudaHostAlloc((void **)&err, 64, cudaHostAllocDefault)) != cudaSuccess) { cudaFree(p); return (int)e; }
		*err = 0;
		unsigned int *expect = NULL;
		if (__atomic_compare_exchange_n(&g_cycle_pool[dev], &expect, p, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) g_cycle_err[dev] = err;
		else { cudaFree(p); cudaFreeHost(err); }                       /* another thread won */
	}
	return 0;
}

/* 1 if a phase wait inside a one-launch cycle on the current device has timed out since the last reset (the batch that
 * was running is suspect); read after synchronising.  reset != 0 clears it. */
extern "C" int gpuhash_cycle_error(int reset)
{
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_cycle_err[dev]) return 0;
	const int v = *g_cycle_err[dev] != 0;
	if (v && reset) *g_cycle_err[dev] = 0;
	return v;
}

extern "C" size_t gpuhash_cycle_workspace_bytes(int max_batches)
{
	if (max_batches < 1) max_batches = 1;
	return ((gh::cycle_workspace_words(max_batches) * sizeof(unsigned int)) + 127) & ~(size_t)127;
}

static unsigned long long cycle_timeout_ns(void)
{
	static unsigned long long v = 0;
	if (!v) { const char *e = getenv("GPUHASH_CYCLE_TIMEOUT_MS"); long ms = e && *e ? atol(e) : 0; v = (unsigned long long)(ms > 0 ? ms : 10000) * 1000000ULL; }
	return v;
}

/* shape + launch shared by every entry point of the one-launch cycle.  total_tiles: from host-side counts (0 = unknown:
 * full persistent grid) */
/* resident CTAs per SM of a cycle kernel variant (the tickets do not need a resident grid to be correct -- it is simply the
 * smallest grid that fills the GPU: fewer span tables to build, fewer ticket atomics at the start) */
template <typename K>
static int cycle_ctas_per_sm(K kernel, int threads, size_t smem, int *cache)
{
	if (*cache <= 0) {
		int n = 0;
		if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess || n < 1) { (void)cudaGetLastError(); n = 2048 / threads / 2; }
		*cache = n;
	}
	return *cache;
}

/* GPUHASH_CYCLE_CTAS_PER_SM / gpuhash_set_cycle_ctas_per_sm: cap on the persistent grid of the cycle kernel (0 = fill the GPU).
 * A cycle whose batches live in pinned HOST memory is bound by the host link, which a fraction of the warps saturates; with
 * one CTA per SM two consecutive cycles are resident together and the link never idles at a cycle boundary. */
static int g_cycle_ctas_cap = -1;
extern "C" void gpuhash_set_cycle_ctas_per_sm(int n) { g_cycle_ctas_cap = n > 0 ? n : 0; }

static int launch_cycle_multi(const gpuhash_geom_t *g, void *table_d, gh::MultiArgs &a, size_t total_tiles, int compact,
		gpuhash_stats_t *stats_d, cudaStream_t s)
{
	if (g_cycle_ctas_cap < 0) { const char *e = getenv("GPUHASH_CYCLE_CTAS_PER_SM"); g_cycle_ctas_cap = e && *e ? atoi(e) : 0; if (g_cycle_ctas_cap < 0) g_cycle_ctas_cap = 0; }
	const size_t sms = (size_t)sm_count_now();
	/* small cycles (a lone 64 K batch is ~1000 tiles) take 2-warp CTAs so that they still reach every SM */
	const unsigned threads = total_tiles && total_tiles <= sms * 16 ? 64u : 256u;
	const size_t wpc = threads / 32;
	const size_t smem = gh::span_table_bytes(a.W, a.num_segs);
	a.timeout_ns = cycle_timeout_ns();
	{	/* requests per delete / insert tile (16, 32 or 64) */
		static int upd_env = -1;
		if (upd_env < 0) { const char *e = getenv("GPUHASH_UPD_TILE"); upd_env = e && (atoi(e) == 16 || atoi(e) == 32 || atoi(e) == 64) ? atoi(e) : 0; }
		/* search-heavy cycles (the caller leaves upd_tile 0) take 16-request update tiles: with 64 the last phase of a 95/5 cycle of
		 * 64 batches is 1.4 tiles per warp (a cycle on its own: 215.8 us with 64, 211.0 with 16); update-heavy cycles (the caller
		 * presets 64: updates >= a quarter of the requests) keep 64 -- a 50/50 cycle loses a quarter with 16 (tickets, waits) */
		if (upd_env) a.upd_tile = (uint32_t)upd_env;
		else if (threads == 64u || a.upd_tile != 64u) a.upd_tile = 16u;
	}
	{
		int dev = 0;
		if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64) {
			if (!__atomic_load_n(&g_cycle_pool[dev], __ATOMIC_ACQUIRE)) (void)gpuhash_init_device();   /* (fails inside a stream capture: */
			a.err_host = g_cycle_err[dev];                                                              /*  then errors stay in ws[2])    */
		}
	}
	gh::Geom gg = to_geom(g);
	gh::Bucket *t = (gh::Bucket *)table_d; gh::Stats *st = (gh::Stats *)stats_d;
	const int variant = (gg.layout == gh::kLayoutPairs ? 0 : 2) + (compact ? 1 : 0);
	static int occ[4][2];                              /* [variant][threads == 256] */
	int *oc = &occ[variant][threads == 256];
	int per_sm;
	switch (variant) {
	case 0:  per_sm = cycle_ctas_per_sm(gh::cycle_multi_kernel<true, false>, (int)threads, smem, oc); break;
	case 1:  per_sm = cycle_ctas_per_sm(gh::cycle_multi_kernel<true, true>, (int)threads, smem, oc); break;
	case 2:  per_sm = cycle_ctas_per_sm(gh::cycle_multi_kernel<false, false>, (int)threads, smem, oc); break;
	default: per_sm = cycle_ctas_per_sm(gh::cycle_multi_kernel<false, true>, (int)threads, smem, oc); break;
	}
	if (g_cycle_ctas_cap > 0 && threads == 256u && per_sm > g_cycle_ctas_cap) per_sm = g_cycle_ctas_cap;
	size_t blocks = sms * (size_t)per_sm;
	if (total_tiles) { const size_t need = (total_tiles + wpc - 1) / wpc; if (need < blocks) blocks = need; }
	if (blocks == 0) blocks = 1;
	switch (variant) {
	case 0:  gh::cycle_multi_kernel<true, false><<<(unsigned)blocks, threads, smem, s>>>(t, gg, st, a); break;
	case 1:  gh::cycle_multi_kernel<true, true><<<(unsigned)blocks, threads, smem, s>>>(t, gg, st, a); break;
	case 2:  gh::cycle_multi_kernel<false, false><<<(unsigned)blocks, threads, smem, s>>>(t, gg, st, a); break;
	default: gh::cycle_multi_kernel<false, true><<<(unsigned)blocks, threads, smem, s>>>(t, gg, st, a); break;
	}
	return (int)cudaGetLastError();
}

static size_t tiles_of_batch(const gpuhash_batch_t *b)
{
	size_t t = 0;
	if (b->n_search) { const size_t head = ((uintptr_t)b->search_in & 15u) ? 1 : 0; const size_t k = (b->n_search - head + gh::kTileReq - 1) / gh::kTileReq; t += k ? k : 1; }
	t += ((size_t)b->n_delete + gh::kTileReq - 1) / gh::kTileReq + ((size_t)b->n_insert + gh::kTileReq - 1) / gh::kTileReq;
	return t;
}

static int batch_ok(const gpuhash_batch_t *b, int compact)
{
	if ((b->n_search && (!b->search_in || !b->search_out)) || (b->n_delete && !b->delete_in) || (b->n_insert && !b->insert_in)) return 0;
	if (((uintptr_t)b->search_in & 7u) || ((uintptr_t)b->search_out & (compact ? 3u : 7u)) || ((uintptr_t)b->delete_in & 3u) || ((uintptr_t)b->insert_in & 3u)) return 0;
	return 1;
}

/* The whole cycle of num_batches workers in ONE launch: per batch search -> delete -> insert (the reference's in-stream
 * order, mega_scheduler.c:392-502), batches unordered against each other (its streams).  batches_d: the descriptor table
 * where the DEVICE reads it (device memory; every CTA reads it, so not pinned host memory); batches_h: the same table on
 * the host, used to size the grid (NULL: full persistent grid).  workspace_d: gpuhash_cycle_workspace_bytes(num_batches)
 * zeroed bytes owned by the caller, one launch at a time.  compact != 0: one result word per search (mega_send.c:411-414). */
extern "C" int gpuhash_cycle_multi_ex(const gpuhash_geom_t *g, void *table_d, const gpuhash_batch_t *batches_h,
		const gpuhash_batch_t *batches_d, int num_batches, int compact, void *workspace_d, gpuhash_stats_t *stats_d, void *stream)
{
	static_assert(sizeof(gpuhash_batch_t) == sizeof(gh::BatchDesc), "descriptor mirror");
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || !table_d || !batches_d || !workspace_d) return -1;
	if (num_batches < 1 || num_batches > gh::kMaxBatches) return -1;
	size_t tiles = 0, n_s = 0, n_u = 0;
	if (batches_h) {
		for (int w = 0; w < num_batches; w++) {
			if (!batch_ok(&batches_h[w], compact)) return -1;
			tiles += tiles_of_batch(&batches_h[w]);
			n_s += batches_h[w].n_search; n_u += (size_t)batches_h[w].n_delete + batches_h[w].n_insert;
		}
		if (tiles == 0) return 0;
	}
	gh::MultiArgs a; memset(&a, 0, sizeof a);
	if (!batches_h || 3 * n_u >= n_s) a.upd_tile = 64;       /* update-heavy (or unknown): see launch_cycle_multi */
	a.descs = (const gh::BatchDesc *)batches_d; a.W = num_batches;
	a.ws = (uint32_t *)workspace_d;
	return launch_cycle_multi(g, table_d, a, tiles, compact, stats_d, (cudaStream_t)stream);
}

/* One launch for a whole scheduler cycle of ONE worker: searches, then deletes, then inserts (flat batch with a
 * host-known count, or -- blk_input_d != NULL -- segments with device-side counts).  Same results as
 * gpuhash_search_ex, gpuhash_delete_ex, gpuhash_insert_*_ex issued in that order on one stream.
 * workspace_d: as gpuhash_cycle_multi_ex (one batch), or NULL = a slot of the per-device pool. */
extern "C" int gpuhash_cycle_ws_ex(const gpuhash_geom_t *g, void *table_d,
		const void *selem_d, size_t n_search, void *out_d,
		const void *delem_d, size_t n_delete,
		const void *ielem_d, size_t n_insert,
		const void *const *blk_input_d, const int *blk_elem_num_d, int num_blks,
		int compact, void *workspace_d, gpuhash_stats_t *stats_d, void *stream)
{
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || !table_d) return -1;
	if (n_search > 0xffffffffULL || n_delete > 0xffffffffULL || n_insert > 0xffffffffULL) return -1;
	if (blk_input_d && (!blk_elem_num_d || num_blks < 1 || n_insert)) return -1;
	gpuhash_batch_t b; memset(&b, 0, sizeof b);
	b.search_in = selem_d; b.search_out = out_d; b.delete_in = delem_d; b.insert_in = ielem_d;
	b.n_search = (uint32_t)n_search; b.n_delete = (uint32_t)n_delete; b.n_insert = (uint32_t)n_insert;
	if (!batch_ok(&b, compact)) return -1;
	cudaStream_t s = (cudaStream_t)stream;
	if (blk_input_d && num_blks > gh::kMaxSegs) {
		/* more segments than the span table holds (the scheduler uses INSERT_BLOCK = 8): the delete and the
		 * count-independent insert launch, stream-ordered -- same results */
		int rc = 0;
		if (n_search && (rc = compact ? gpuhash_search_compact_ex(g, selem_d, out_d, table_d, n_search, stats_d, stream)
		                              : gpuhash_search_ex(g, selem_d, out_d, table_d, n_search, stats_d, stream)) != 0) return rc;
		if (n_delete && (rc = gpuhash_delete_ex(g, delem_d, table_d, n_delete, stats_d, 0, stream)) != 0) return rc;
		return gpuhash_insert_ex(g, table_d, blk_input_d, blk_elem_num_d, num_blks, stats_d, 0, stream);
	}
	size_t tiles = tiles_of_batch(&b);
	if (tiles == 0 && !blk_input_d) return 0;
	gh::MultiArgs a; memset(&a, 0, sizeof a);
	a.descs = nullptr; a.W = 1; memcpy(&a.d0, &b, sizeof b);
	if (blk_input_d) {
		a.seg_ptrs = (const uint32_t *const *)blk_input_d; a.seg_counts = blk_elem_num_d; a.num_segs = num_blks;
		/* segment sizes live in device memory (mega_scheduler.c:493-494): assume the server's bound per segment
		 * (batch_max_insert_job = 4096, mega.c:143) for the grid; the kernel itself reads the real counts */
		tiles += (size_t)num_blks * (4096 / gh::kTileReq);
	}
	if (blk_input_d || 3 * (n_delete + n_insert) >= n_search) a.upd_tile = 64;      /* update-heavy (or sizes only known on the device) */
	if (workspace_d) a.ws = (uint32_t *)workspace_d;
	else {
		int dev = 0;
		if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
		if (!__atomic_load_n(&g_cycle_pool[dev], __ATOMIC_ACQUIRE)) {      /* first use on this device (not allowed while a stream is   */
			int rc = gpuhash_init_device();                                /* capturing: gpuhash_index_create and the bench loops call   */
			if (rc) return rc;                                             /* this up front)                                             */
		}
		const unsigned slot = __atomic_fetch_add(&g_cycle_next[dev], 1u, __ATOMIC_RELAXED) % GH_CYCLE_SLOTS;
		a.ws = g_cycle_pool[dev] + (size_t)GH_POOL_WS_WORDS * slot;
	}
	return launch_cycle_multi(g, table_d, a, tiles, compact, stats_d, s);
}

extern "C" int gpuhash_cycle_ex(const gpuhash_geom_t *g, void *table_d,
		const void *selem_d, size_t n_search, void *out_d,
		const void *delem_d, size_t n_delete,
		const void *ielem_d, size_t n_insert,
		const void *const *blk_input_d, const int *blk_elem_num_d, int num_blks,
		gpuhash_stats_t *stats_d, void *stream)
{
	return gpuhash_cycle_ws_ex(g, table_d, selem_d, n_search, out_d, delem_d, n_delete, ielem_d, n_insert,
			blk_input_d, blk_elem_num_d, num_blks, 0, NULL, stats_d, stream);
}

/* ------------------------------------------------------------------ legacy C ABI */

/* reference gpu_hash.cu:482-518.  num_thread / threads_per_blk described the reference's own
 * grid; they are validated the way a caller could observe (threads_per_blk <= 1024) and not used. */
extern "C" void gpu_hash_search(selem_t *in, loc_t *out, bucket_t *hash_table,
		int num_elem, int num_thread, int threads_per_blk, cudaStream_t stream)
{
	assert(num_elem >= 0);
	assert(threads_per_blk > 0 && threads_per_blk <= 1024);
	assert(num_thread > 0);
	(void)num_thread; (void)threads_per_blk;
	int rc = gpuhash_search_ex(&g_default_geom, in, out, hash_table, (size_t)num_elem, NULL, (void *)stream);
	assert(rc >= 0); (void)rc;       /* CUDA errors surface at the caller's next CUDA_SAFE_CALL, as before */
}

/* reference gpu_hash.cu:521-556 */
extern "C" void gpu_hash_insert(bucket_t *hash_table, ielem_t **blk_input, int *blk_elem_num,
		int num_blks, cudaStream_t stream)
{
	assert(num_blks >= 0);
	int rc = gpuhash_insert_ex(&g_default_geom, hash_table, (const void *const *)blk_input, blk_elem_num,
			num_blks, NULL, 0, (void *)stream);
	assert(rc >= 0); (void)rc;
}

/* reference gpu_hash.cu:558-593 */
extern "C" void gpu_hash_delete(delem_t *in, bucket_t *hash_table, int num_elem, int num_thread,
		int threads_per_blk, cudaStream_t stream)
{
	assert(num_elem >= 0);
	assert(threads_per_blk > 0 && threads_per_blk <= 1024);
	(void)num_thread; (void)threads_per_blk;
	int rc = gpuhash_delete_ex(&g_default_geom, in, hash_table, (size_t)num_elem, NULL, 0, (void *)stream);
	assert(rc >= 0); (void)rc;
}

/* reference libgpuhash.h:53-62 (declared, never defined there): delete batch, then insert batch,
 * stream-ordered so the result equals gpu_hash_delete followed by gpu_hash_insert. */
extern "C" void gpu_delete_insert(bucket_t *hash_table, delem_t *delete_in, uint32_t num_delete_job,
		ielem_t **insert_blk_input, int *insert_blk_elem_num, int num_insert_blks,
		uint32_t num_delete_thread, uint32_t threads_per_blk, cudaStream_t stream)
{
	(void)num_delete_thread; (void)threads_per_blk;
	int rc;
	if (num_insert_blks > 0)
		rc = gpuhash_cycle_ex(&g_default_geom, hash_table, NULL, 0, NULL, delete_in, (size_t)num_delete_job, NULL, 0,
				(const void *const *)insert_blk_input, insert_blk_elem_num, num_insert_blks, NULL, (void *)stream);
	else
		rc = gpuhash_delete_ex(&g_default_geom, delete_in, hash_table, (size_t)num_delete_job, NULL, 0, (void *)stream);
	assert(rc >= 0); (void)rc;
}

/* ------------------------------------------------------------------ plumbing */

extern "C" int gpuhash_device_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
extern "C" int gpuhash_set_device(int dev) { return (int)cudaSetDevice(dev); }

extern "C" int gpuhash_device_info(int dev, int *sm_count, int *l2_bytes, size_t *free_bytes, size_t *total_bytes)
{
	cudaError_t e;
	if (sm_count && (e = cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return (int)e;
	if (l2_bytes && (e = cudaDeviceGetAttribute(l2_bytes, cudaDevAttrL2CacheSize, dev)) != cudaSuccess) return (int)e;
	if (free_bytes || total_bytes) {
		size_t f = 0, t = 0;
		if ((e = cudaMemGetInfo(&f, &t)) != cudaSuccess) return (int)e;
		if (free_bytes) *free_bytes = f;
		if (total_bytes) *total_bytes = t;
	}
	return 0;
}

extern "C" void *gpuhash_dev_alloc(size_t bytes) { void *p = NULL; return cudaMalloc(&p, bytes) == cudaSuccess ? p : NULL; }
extern "C" int gpuhash_dev_free(void *p) { return (int)cudaFree(p); }
extern "C" int gpuhash_dev_memset(void *p, int v, size_t bytes, void *stream)
{
	return (int)cudaMemsetAsync(p, v, bytes, (cudaStream_t)stream);
}
extern "C" int gpuhash_h2d(void *dst_d, const void *src_h, size_t bytes, void *stream)
{
	return (int)cudaMemcpyAsync(dst_d, src_h, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
}
extern "C" int gpuhash_d2h(void *dst_h, const void *src_d, size_t bytes, void *stream)
{
	return (int)cudaMemcpyAsync(dst_h, src_d, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}
extern "C" void *gpuhash_host_alloc(size_t bytes) { void *p = NULL; return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : NULL; }
extern "C" int gpuhash_host_free(void *p) { return (int)cudaFreeHost(p); }
extern "C" void *gpuhash_stream_create(void) { cudaStream_t s = NULL; return cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess ? (void *)s : NULL; }
extern "C" int gpuhash_stream_destroy(void *s) { return (int)cudaStreamDestroy((cudaStream_t)s); }
extern "C" int gpuhash_stream_sync(void *s) { return (int)cudaStreamSynchronize((cudaStream_t)s); }
extern "C" int gpuhash_device_sync(void) { return (int)cudaDeviceSynchronize(); }
extern "C" void *gpuhash_event_create(void) { cudaEvent_t e = NULL; return cudaEventCreate(&e) == cudaSuccess ? (void *)e : NULL; }
extern "C" int gpuhash_event_destroy(void *e) { return (int)cudaEventDestroy((cudaEvent_t)e); }
extern "C" int gpuhash_event_record(void *e, void *s) { return (int)cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s); }
extern "C" int gpuhash_event_elapsed_ms(void *a, void *b, float *ms)
{
	cudaError_t e = cudaEventSynchronize((cudaEvent_t)b);
	if (e != cudaSuccess) return (int)e;
	return (int)cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b);
}
/* DRAM->L2 fetch granularity hint of the current device (cudaLimitMaxL2FetchGranularity: 32, 64 or 128 bytes).
 * A hash probe wants one 32 B sector per random access; see DESIGN.md "L2 fetch granularity". */
extern "C" int gpuhash_set_l2_fetch_granularity(int bytes)
{
	return (int)cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)bytes);
}
extern "C" int gpuhash_get_l2_fetch_granularity(void)
{
	size_t v = 0;
	return cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity) == cudaSuccess ? (int)v : -1;
}

extern "C" const char *gpuhash_error_string(int err) { return err < 0 ? "bad argument" : cudaGetErrorString((cudaError_t)err); }
extern "C" const char *gpuhash_build_info(void)
{
	return "megakv_b200 libgpuhash: sm_100a, built " __DATE__ " " __TIME__;
}

/* ------------------------------------------------------------------ roofline probe */

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

/* Every thread reads kIlp independent random sectors per round; addresses come from a hash of the
 * thread index, so the only memory traffic is the gather itself.  The xor of everything read is
 * stored only if it equals an impossible value, which keeps the loads alive. */
template <int kIlp, int kMode>
__global__ void __launch_bounds__(256)
gather_kernel(const uint32_t *__restrict__ table, uint64_t unit_mask, size_t n, uint32_t seed, uint32_t *sink)
{
	uint32_t acc = 0;
	const size_t stride = (size_t)gridDim.x * blockDim.x * kIlp;
	for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kIlp; i < n; i += stride) {
		gh::Row r[kIlp]; gh::Row r2[kIlp];
#pragma unroll
		for (int k = 0; k < kIlp; k++) {
			uint64_t u = mix64((uint64_t)(i + k) * 0x9E3779B97F4A7C15ULL + seed) & unit_mask;
			const uint32_t *p = kMode == 1 ? table + u * 8 : table + u * 16;
			r[k] = gh::ld_row_ro(p);
			if (kMode == 2) r2[k] = gh::ld_row_ro(p + 8);
		}
#pragma unroll
		for (int k = 0; k < kIlp; k++) {
#pragma unroll
			for (int l = 0; l < 8; l++) { acc ^= r[k].w[l]; if (kMode == 2) acc ^= r2[k].w[l]; }
		}
	}
	if (acc == 0xDEADBEEFu && seed == 0x12345u) *sink = acc;
}

template <int kIlp>
void launch_gather(int mode, const uint32_t *t, uint64_t mask, size_t n, uint32_t seed, uint32_t *sink,
		unsigned blocks, cudaStream_t s)
{
	if (mode == 0)      gather_kernel<kIlp, 0><<<blocks, 256, 0, s>>>(t, mask, n, seed, sink);
	else if (mode == 1) gather_kernel<kIlp, 1><<<blocks, 256, 0, s>>>(t, mask, n, seed, sink);
	else                gather_kernel<kIlp, 2><<<blocks, 256, 0, s>>>(t, mask, n, seed, sink);
}

}  // namespace

extern "C" int gpuhash_roofline_gather(const void *table_d, size_t table_bytes, size_t n, int mode,
		int loads_per_thread, int iters, float *best_ms, void *stream)
{
	if (!table_d || !best_ms || table_bytes < 64 || (table_bytes & (table_bytes - 1)) || mode < 0 || mode > 2) return -1;
	cudaStream_t s = (cudaStream_t)stream;
	uint64_t mask = (mode == 1 ? table_bytes / 32 : table_bytes / 64) - 1;
	uint32_t *sink = NULL;
	cudaError_t e = cudaMalloc(&sink, 4);
	if (e != cudaSuccess) return (int)e;
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	int ilp = loads_per_thread >= 8 ? 8 : loads_per_thread >= 4 ? 4 : loads_per_thread >= 2 ? 2 : 1;
	size_t blocks = (n + (size_t)256 * ilp - 1) / ((size_t)256 * ilp);
	if (blocks > 0x7fffffffULL) blocks = 0x7fffffffULL;
	float best = 1e30f;
	for (int it = 0; it < iters + 1; it++) {                 /* first launch is warm-up */
		cudaEventRecord(a, s);
		uint32_t seed = 1000u + (uint32_t)it;
		const uint32_t *t = (const uint32_t *)table_d;
		if (ilp == 8)      launch_gather<8>(mode, t, mask, n, seed, sink, (unsigned)blocks, s);
		else if (ilp == 4) launch_gather<4>(mode, t, mask, n, seed, sink, (unsigned)blocks, s);
		else if (ilp == 2) launch_gather<2>(mode, t, mask, n, seed, sink, (unsigned)blocks, s);
		else               launch_gather<1>(mode, t, mask, n, seed, sink, (unsigned)blocks, s);
		cudaEventRecord(b, s);
		if ((e = cudaEventSynchronize(b)) != cudaSuccess) break;
		float ms = 0; cudaEventElapsedTime(&ms, a, b);
		if (it > 0 && ms < best) best = ms;
	}
	cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
	*best_ms = best;
	return (int)e;
}
