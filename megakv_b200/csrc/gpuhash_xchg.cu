/*
 * gpuhash_xchg.cu -- the sharded index as ONE kernel per scheduler cycle: routing, lookup and un-routing fused and
 * software-pipelined over NVLink peer memory (BASELINE.json north_star (d)).
 *
 * The reference is single-GPU (src/mega.c:410); what it does per scheduler cycle -- walk every worker's batch,
 * search -> delete -> insert, synchronise once (src/mega_scheduler.c:393-504) -- is here one *exchange* per GPU.
 * gpuhash_shard.cu runs an exchange as scatter / serve / gather kernels behind flag waits; they do not overlap
 * (each fills the GPU), so a routed search paid 13 us per million for the routing next to the lookup's 48 and a
 * GPU of the sharded index reached 15-16 Gops/s against 20 alone.  Here launch j of a rank does, in ONE kernel,
 *
 *      scatter  of exchange j      its requests, sorted by owner, into the owners' inboxes      (peer stores)
 *      serve    of exchange j-1    everything the peers put into its inbox one launch ago:
 *                                  searches (results straight into the origins' staging areas,   (peer stores)
 *                                  then deletes, then inserts (the reference's in-stream order)
 *      gather   of exchange j-2    results the owners staged one launch ago, back into request order
 *
 * Nothing a launch reads from a peer was produced later than that peer's PREVIOUS launch, so no thread ever waits for
 * another GPU inside a kernel: one stream memory operation in front of launch j (flag[s] >= j-1 for every peer s)
 * is the whole inter-GPU synchronisation.  Inside the launch the three kinds of work are independent; warps take
 * tiles by an atomic ticket from an index space that INTERLEAVES them -- per group of six tickets: one scatter tile
 * (256 requests), four lookup tiles (64 requests each), one gather tile (256 results) -- so the streaming traffic of
 * the routing (NVLink stores, coalesced reads) rides in the shadow of the lookups' random line fills instead of in
 * front of and behind them.  Delete and insert tiles come last in ticket order and wait (device-scope counters) for
 * the lookups of this launch, as a worker's insert follows its search in the reference.
 *
 * Buffers are triple-buffered by exchange number (slot = e mod 3): a peer may already be scattering exchange j+1
 * into slot (j+1) mod 3 of my inbox while I still serve exchange j-1 from slot (j-1) mod 3; slot e mod 3 is written
 * again by launch e+3 at the earliest, which waits for the flags of launch e+2, raised after everyone's use of e.
 *
 * A warp tile of the scatter: 256 requests (16 B per lane and load), owner = top bits of bucket 1 (== of bucket 2 and of
 * every eviction target, gpu_hash.h:67-69), rank inside (tile, owner) by ballots, one global atomic per (tile, owner)
 * reserves the run in the owner's region, the tile is sorted through shared memory and every run leaves as contiguous
 * stores.  The map the gather needs: one byte per request (its place in the sorted tile) + 64 B per tile (run starts,
 * lengths).  A warp tile of the gather reads the runs back contiguously and writes the results in request order.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>
#include "gpuhash_ex.h"
#include "gpuhash_kernels.cuh"

namespace {

constexpr int kMaxShards = 8;
constexpr uint32_t kRTile = 256;           /* requests per scatter / gather tile (one warp) */
constexpr uint32_t kPTile = gh::kTileReq;  /* requests per lookup tile: 64 */
constexpr uint32_t kUTile = 64;            /* requests per delete / insert tile */
constexpr uint32_t kGroup = 6;             /* tickets per interleave group: 1 scatter + 4 lookup + 1 gather */
constexpr uint32_t kMaxClaim = 6;

/* byte offsets inside a rank's arena; identical on every rank (peers address each other's arenas with them) */
struct XLayout {
	/* written by peers */
	size_t inbox_s;        /* [3][G][cap_s] selem_t      requests of source s for me, exchange slot e % 3 */
	size_t inbox_d;        /* [3][G][cap_u] delem_t */
	size_t inbox_i;        /* [3][G][cap_u] ielem_t */
	size_t stage;          /* [3][G][cap_s] loc_t[2]     results of owner d for my requests */
	size_t cnt;            /* [3][3][8] u32              how many requests source s put into each inbox (kind-major) */
	size_t flag;           /* [8] u32                    flag[s] = number of the last launch of rank s whose stores are complete */
	/* local */
	size_t pos;            /* [3][cap_s] u8              place of request i in its sorted tile */
	size_t meta;           /* [3][cap_s / 256][16] u32   per tile: run start in the owner's region [8], run length [8] */
	size_t counts;         /* [3][3][8] u32              slots handed out per (kind, owner); zero when the slot is free */
	size_t ws;             /* [32] u32                   ticket, CTAs done, warps out of the interleaved region, tiles done */
	size_t err;            /* u32                        sticky: a wait inside a kernel timed out */
	size_t total;
	uint32_t cap_s, cap_u;
};

enum { kWsTicket = 0, kWsCtas = 8, kWsRegion = 16, kWsYDone = 24, kWsUDone = 32, kWsWords = 40 };   /* one 32 B sector each */

struct XArgs {
	gh::Bucket *table; gh::Geom g; gh::Stats *st;
	char *peer[kMaxShards];                /* every rank's arena as mapped into this process; peer[rank] is my own */
	XLayout L;
	int G, rank;
	uint32_t seq;                          /* number of this launch, 1.. */
	uint32_t hash_mask_total; int shift;
	const uint2 *s_in; const uint32_t *d_in; const uint32_t *i_in;     /* exchange seq: my requests (device or pinned host) */
	uint32_t s_n, d_n, i_n;
	int do_serve;                          /* exchange seq-1 exists */
	uint2 *g_out; uint32_t g_n;            /* exchange seq-2: where its results go */
	unsigned long long timeout_ns;
};

__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p)
{
	uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
	uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

/* run (owner) of position q in a sorted tile; off[1..7] = run starts, all in shared memory */
__device__ __forceinline__ int run_of(const uint32_t *off, uint32_t q)
{
	return (int)(q >= off[1]) + (int)(q >= off[2]) + (int)(q >= off[3]) + (int)(q >= off[4])
	     + (int)(q >= off[5]) + (int)(q >= off[6]) + (int)(q >= off[7]);
}

struct Plan {                              /* built by every CTA in shared memory */
	uint32_t xfirst[4];                    /* scatter tiles: search | delete | insert */
	uint32_t ycnt[kMaxShards], yfirst[kMaxShards + 1];      /* lookup tiles per source */
	uint32_t ucnt[kMaxShards], ufirst[kMaxShards + 1];      /* delete tiles per source */
	uint32_t vcnt[kMaxShards], vfirst[kMaxShards + 1];      /* insert tiles per source */
	uint32_t nZ, inter, total;
};

struct Tile { int kind; uint32_t idx; };   /* kind: 0 none, 1 scatter, 2 lookup, 3 gather, 4 delete, 5 insert */

__device__ __forceinline__ Tile decode(const Plan &P, uint32_t t)
{
	Tile r; r.kind = 0; r.idx = 0;
	if (t >= P.total) return r;
	if (t < P.inter) {
		const uint32_t grp = t / kGroup, k = t - grp * kGroup;
		if (k == 0)               { if (grp < P.xfirst[3]) { r.kind = 1; r.idx = grp; } }
		else if (k == kGroup - 1) { if (grp < P.nZ) { r.kind = 3; r.idx = grp; } }
		else { const uint32_t y = 4 * grp + k - 1; if (y < P.yfirst[kMaxShards]) { r.kind = 2; r.idx = y; } }
		return r;
	}
	const uint32_t u = t - P.inter;
	if (u < P.ufirst[kMaxShards]) { r.kind = 4; r.idx = u; }
	else { r.kind = 5; r.idx = u - P.ufirst[kMaxShards]; }
	return r;
}

__device__ __forceinline__ int source_of(const uint32_t *first, uint32_t idx)      /* first[s] <= idx < first[s+1] */
{
	int s = 0;
	while (s < kMaxShards - 1 && idx >= first[s + 1]) s++;
	return s;
}

/* whole warp: wait until *c >= target */
__device__ __forceinline__ void counter_wait(const uint32_t *c, uint32_t target, const XArgs &a, unsigned lane)
{
	if (target == 0) return;
	if (lane == 0 && ld_acquire_gpu(c) < target) {
		unsigned long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
		unsigned ns = 128u;
		for (;;) {
			__nanosleep(ns);
			if (ld_acquire_gpu(c) >= target) break;
			if (ns < 2048u) ns <<= 1;
			unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
			if (t1 - t0 > a.timeout_ns) { atomicExch((uint32_t *)(a.peer[a.rank] + a.L.err), 1u); break; }
		}
	}
	__syncwarp();
}

/* ---- scatter: one warp, one tile of 256 requests of kWords words ---- */
template <int kWords>
__device__ __forceinline__ void scatter_tile(const XArgs &a, const uint32_t *in, uint32_t n, uint32_t tile, int kind, uint32_t slot,
		uint32_t *stage /* this warp's kRTile * 3 words */, uint32_t *runs /* this warp's 32 words */, unsigned lane)
{
	const uint32_t t0 = tile * kRTile, tile_n = min(kRTile, n - t0);
	const int G = a.G;
	uint32_t w[8][kWords], k_of[8];
	bool live[8];
	if (kWords == 2) {                                     /* requests 64*it + 2*lane + {0, 1}: 16 B per lane, 512 B per warp load */
		const uint2 *p = (const uint2 *)in + t0;
		const bool vec = ((uintptr_t)in & 15u) == 0;
#pragma unroll
		for (int it = 0; it < 4; it++) {
			const uint32_t k = 64 * it + 2 * lane;
			k_of[2 * it] = k; k_of[2 * it + 1] = k + 1;
			live[2 * it] = k < tile_n; live[2 * it + 1] = k + 1 < tile_n;
			uint4 v = make_uint4(0u, 0u, 0u, 0u);
			if (vec && live[2 * it + 1]) {
				asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + k));
			} else {
				if (live[2 * it]) { const uint2 q = gh::ld_stream_u2(p + k); v.x = q.x; v.y = q.y; }
				if (live[2 * it + 1]) { const uint2 q = gh::ld_stream_u2(p + k + 1); v.z = q.x; v.w = q.y; }
			}
			w[2 * it][0] = v.x; w[2 * it][1] = v.y; w[2 * it + 1][0] = v.z; w[2 * it + 1][1] = v.w;
		}
	} else {                                               /* 12-byte records: the raw tile through shared memory, word by word */
		const uint32_t *p = in + (size_t)3 * t0;
		for (uint32_t q = lane; q < 3 * tile_n; q += 32) stage[q] = gh::ld_stream_u32(p + q);
		__syncwarp();
#pragma unroll
		for (int r = 0; r < 8; r++) {
			const uint32_t k = 32 * r + lane;
			k_of[r] = k; live[r] = k < tile_n;
#pragma unroll
			for (int j = 0; j < kWords; j++) w[r][j] = live[r] ? stage[3 * k + j] : 0u;      /* stride 3 words: conflict-free */
		}
		__syncwarp();                                      /* everyone has its records before the sorted tile overwrites them */
	}
	/* rank inside (tile, owner): ballots, the running counts are warp-uniform registers */
	uint32_t run[kMaxShards], rank[8], d[8];
#pragma unroll
	for (int o = 0; o < kMaxShards; o++) run[o] = 0;
	const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
	for (int r = 0; r < 8; r++) {
		d[r] = live[r] ? (w[r][1] & a.hash_mask_total) >> a.shift : 0xffu;      /* owner = top bits of bucket 1 */
		rank[r] = 0;
#pragma unroll
		for (int o = 0; o < kMaxShards; o++) {
			if (o < G) {
				const uint32_t b = __ballot_sync(0xffffffffu, d[r] == (uint32_t)o);
				if (d[r] == (uint32_t)o) rank[r] = run[o] + __popc(b & lt_mask);
				run[o] += __popc(b);
			}
		}
	}
	/* lane o: reserve owner o's run in its region, publish run start / length for the gather */
	uint32_t mine = 0, off_mine = 0;
#pragma unroll
	for (int o = 0; o < kMaxShards; o++) { if ((int)lane == o) mine = run[o]; if ((int)lane > o) off_mine += run[o]; }
	uint32_t base = 0;
	uint32_t *counts = (uint32_t *)(a.peer[a.rank] + a.L.counts) + (slot * 3 + kind) * 8;
	if (lane < (unsigned)G && mine) base = atomicAdd(counts + lane, mine);
	if (lane <= kMaxShards) runs[lane] = lane < kMaxShards ? off_mine : tile_n;      /* off[0..8] */
	if (lane < kMaxShards) runs[16 + lane] = base - off_mine;                        /* region index of sorted position q: delta[o] + q */
	if (kind == 0 && lane < kMaxShards) {
		uint32_t *meta = (uint32_t *)(a.peer[a.rank] + a.L.meta) + ((size_t)slot * (a.L.cap_s / kRTile) + tile) * 16;
		meta[lane] = base; meta[8 + lane] = mine;
	}
	__syncwarp();
	/* the sorted tile */
	uint8_t *pos = (uint8_t *)(a.peer[a.rank] + a.L.pos) + (size_t)slot * a.L.cap_s + t0;
	uint32_t pp[8];
#pragma unroll
	for (int r = 0; r < 8; r++) {
		pp[r] = 0;
		if (live[r]) {
			pp[r] = runs[d[r]] + rank[r];
#pragma unroll
			for (int j = 0; j < kWords; j++) stage[pp[r] * kWords + j] = w[r][j];
		}
	}
	if (kind == 0) {
#pragma unroll
		for (int it = 0; it < 4; it++) {                   /* this lane's requests 2*it, 2*it+1 are neighbours: one 2-byte store */
			if (live[2 * it + 1]) *(uint16_t *)(pos + k_of[2 * it]) = (uint16_t)(pp[2 * it] | (pp[2 * it + 1] << 8));
			else if (live[2 * it]) pos[k_of[2 * it]] = (uint8_t)pp[2 * it];
		}
	}
	__syncwarp();
	/* runs out: neighbouring lanes share a run -> contiguous stores, local or over NVLink */
	if (kWords == 2) {
		const size_t region = ((size_t)slot * G + a.rank) * a.L.cap_s;
		for (uint32_t q = lane; q < tile_n; q += 32) {
			const int o = run_of(runs, q);
			const uint32_t at = runs[16 + o] + q;              /* 32-bit wrap-around intended: delta may be "negative" */
			uint2 *dst = (uint2 *)(a.peer[o] + a.L.inbox_s) + region + at;
			*dst = make_uint2(stage[2 * q], stage[2 * q + 1]);
		}
	} else {
		const size_t region = ((size_t)slot * G + a.rank) * a.L.cap_u;
		const size_t box = kind == 1 ? a.L.inbox_d : a.L.inbox_i;
		for (uint32_t wq = lane; wq < 3 * tile_n; wq += 32) {
			const uint32_t q = wq / 3, j = wq - 3 * q;
			const int o = run_of(runs, q);
			const uint32_t at = runs[16 + o] + q;              /* 32-bit wrap-around intended: delta may be "negative" */
			uint32_t *dst = (uint32_t *)(a.peer[o] + box) + 3 * (region + at) + j;
			*dst = stage[wq];
		}
	}
	__syncwarp();                                          /* stage / runs are free again */
}

/* ---- gather: one warp, one tile of 256 results back into request order ---- */
__device__ __forceinline__ void gather_tile(const XArgs &a, uint32_t tile, uint32_t slot, uint32_t *stage, uint32_t *runs, unsigned lane)
{
	const uint32_t t0 = tile * kRTile, tile_n = min(kRTile, a.g_n - t0);
	const int G = a.G;
	const uint32_t *meta = (const uint32_t *)(a.peer[a.rank] + a.L.meta) + ((size_t)slot * (a.L.cap_s / kRTile) + tile) * 16;
	uint32_t base = 0, len = 0;
	if (lane < kMaxShards) { base = meta[lane]; len = meta[8 + lane]; }
	uint32_t inc = len;
#pragma unroll
	for (int dd = 1; dd < kMaxShards; dd <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, dd); if ((int)lane >= dd) inc += o; }
	if (lane < kMaxShards) { runs[lane] = inc - len; runs[16 + lane] = base - (inc - len); }
	if (lane == kMaxShards) runs[kMaxShards] = tile_n;
	__syncwarp();
	uint2 *st2 = (uint2 *)stage;
	const uint2 *mine = (const uint2 *)(a.peer[a.rank] + a.L.stage) + (size_t)slot * G * a.L.cap_s;
	for (uint32_t q = lane; q < tile_n; q += 32) {
		const int o = run_of(runs, q);
		uint2 v;
		const uint32_t at = runs[16 + o] + q;
		const uint2 *src = mine + (size_t)o * a.L.cap_s + at;
		asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(src));
		st2[q] = v;
	}
	__syncwarp();
	const uint8_t *pos = (const uint8_t *)(a.peer[a.rank] + a.L.pos) + (size_t)slot * a.L.cap_s + t0;
	uint2 *out = a.g_out + t0;
	const bool vec = ((uintptr_t)a.g_out & 15u) == 0;
#pragma unroll
	for (int it = 0; it < 4; it++) {
		const uint32_t k = 64 * it + 2 * lane;
		if (k + 1 < tile_n) {
			const uint32_t p2 = *(const uint16_t *)(pos + k);
			const uint2 r0 = st2[p2 & 255u], r1 = st2[p2 >> 8];
			if (vec) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + k), "r"(r0.x), "r"(r0.y), "r"(r1.x), "r"(r1.y) : "memory");
			else { gh::st_stream_u2(out + k, r0); gh::st_stream_u2(out + k + 1, r1); }
		} else if (k < tile_n) {
			gh::st_stream_u2(out + k, st2[pos[k]]);
		}
	}
	__syncwarp();
}

#ifndef GH_XCHG_MIN_CTAS
#define GH_XCHG_MIN_CTAS 2
#endif
template <bool kPairs>
__global__ void __launch_bounds__(256, GH_XCHG_MIN_CTAS)
xchg_step_kernel(XArgs a)
{
	__shared__ Plan P;
	__shared__ __align__(16) uint32_t stage_s[8][kRTile * 3];
	__shared__ uint32_t runs_s[8][32];
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const int G = a.G;
	const uint32_t slot_x = a.seq % 3u, slot_y = (a.seq + 2u) % 3u, slot_z = (a.seq + 1u) % 3u;
	char *me = a.peer[a.rank];
	uint32_t *ws = (uint32_t *)(me + a.L.ws);

	if (threadIdx.x == 0) {
		P.xfirst[0] = 0;
		P.xfirst[1] = (a.s_n + kRTile - 1) / kRTile;
		P.xfirst[2] = P.xfirst[1] + (a.d_n + kRTile - 1) / kRTile;
		P.xfirst[3] = P.xfirst[2] + (a.i_n + kRTile - 1) / kRTile;
		const uint32_t *cnt = (const uint32_t *)(me + a.L.cnt) + slot_y * 24;     /* written by the peers one launch ago */
		uint32_t y = 0, u = 0, v = 0;
		for (int s = 0; s < kMaxShards; s++) {
			const bool on = a.do_serve && s < G;
			const uint32_t cs = on ? cnt[s] : 0u, cd = on ? cnt[8 + s] : 0u, ci = on ? cnt[16 + s] : 0u;
			P.ycnt[s] = cs; P.yfirst[s] = y; y += (cs + kPTile - 1) / kPTile;
			P.ucnt[s] = cd; P.ufirst[s] = u; u += (cd + kUTile - 1) / kUTile;
			P.vcnt[s] = ci; P.vfirst[s] = v; v += (ci + kUTile - 1) / kUTile;
		}
		P.yfirst[kMaxShards] = y; P.ufirst[kMaxShards] = u; P.vfirst[kMaxShards] = v;
		P.nZ = (a.g_n + kRTile - 1) / kRTile;
		const uint32_t groups = max(max(P.xfirst[3], (y + 3) / 4), P.nZ);
		P.inter = kGroup * groups;
		P.total = P.inter + u + v;
	}
	__syncthreads();

	const uint32_t total = P.total, inter = P.inter;
	const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
	uint32_t h1 = 0, h2 = 0;
	uint32_t pend_y = 0, pend_u = 0;                      /* tiles this warp finished and has not published yet */
	bool in_region = true;

	auto publish = [&]() {                                /* device scope: what the delete / insert tiles of this launch wait for */
		if (pend_y == 0 && pend_u == 0) return;           /* warp-uniform */
		__syncwarp();
		if (lane == 0) {
			__threadfence();
			if (pend_y) atomicAdd(ws + kWsYDone, pend_y);
			if (pend_u) atomicAdd(ws + kWsUDone, pend_u);
		}
		pend_y = 0; pend_u = 0;
	};
	/* Leaving the interleaved region: everything this warp stored for peers (scattered requests, results) is complete at
	 * system scope before it counts itself; the last warp to leave hands the peers my counts and raises my flag. */
	auto leave_region = [&]() {
		publish();
		__syncwarp();
		uint32_t last = 0;
		if (lane == 0) {
			__threadfence_system();
			last = atomicAdd(ws + kWsRegion, 1u) == nwarps - 1 ? 1u : 0u;
		}
		last = __shfl_sync(0xffffffffu, last, 0);
		if (last && lane < (unsigned)G) {
			__threadfence();
			const uint32_t *counts = (const uint32_t *)(me + a.L.counts) + slot_x * 24;
			volatile uint32_t *pc = (volatile uint32_t *)(a.peer[lane] + a.L.cnt) + slot_x * 24;
			pc[a.rank] = ld_relaxed_gpu(counts + lane);
			pc[8 + a.rank] = ld_relaxed_gpu(counts + 8 + lane);
			pc[16 + a.rank] = ld_relaxed_gpu(counts + 16 + lane);
			__threadfence_system();
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)(a.peer[lane] + a.L.flag) + a.rank), "r"(a.seq) : "memory");
		}
		in_region = false;
	};

	auto claim_size = [&](uint32_t seen) -> uint32_t {
		const uint32_t rem = total > seen ? total - seen : 0u;
		return min(kMaxClaim, max(1u, rem / (6u * nwarps)));
	};
	auto claim_issue = [&](uint32_t k) -> uint32_t { return lane == 0 ? atomicAdd(ws + kWsTicket, k) : 0u; };
	/* the requests of a lookup tile (this lane's 16 B): in flight while the tile before it is worked on */
	auto y_in = [&](int s) -> const uint2 * { return (const uint2 *)(me + a.L.inbox_s) + ((size_t)slot_y * G + s) * a.L.cap_s; };
	auto prefetch = [&](const Tile &tl) -> uint4 {
		if (tl.kind != 2) return make_uint4(0u, 0u, 0u, 0u);
		const int s = source_of(P.yfirst, tl.idx);
		const uint32_t r0 = (tl.idx - P.yfirst[s]) * kPTile;
		return gh::warp_tile_load<false>(y_in(s) + r0, min(kPTile, P.ycnt[s] - r0), lane);
	};

	uint32_t k_cur = claim_size(0), k_nxt = k_cur;
	const uint32_t raw_a = claim_issue(k_cur), raw_b = claim_issue(k_nxt);
	uint32_t c0 = __shfl_sync(0xffffffffu, raw_a, 0), c1 = c0 + k_cur;
	uint32_t n0 = __shfl_sync(0xffffffffu, raw_b, 0), n1 = n0 + k_nxt;
	uint32_t seen = n1, raw_nn = 0, k_nn = 1;
	uint32_t t = c0;
	Tile cur = decode(P, t);
	uint4 v = prefetch(cur);
	while (t < total) {
		if (t == c0) { k_nn = claim_size(seen); raw_nn = claim_issue(k_nn); }
		if (in_region && t >= inter) leave_region();
		const bool last_of_chunk = t + 1 >= c1;
		const uint32_t tn = last_of_chunk ? n0 : t + 1;
		const Tile nxt = decode(P, tn);
		const uint4 vn = prefetch(nxt);
		if (cur.kind == 1) {
			const uint32_t x = cur.idx;
			if (x < P.xfirst[1]) {
				scatter_tile<2>(a, (const uint32_t *)a.s_in, a.s_n, x, 0, slot_x, stage_s[wid], runs_s[wid], lane);
			} else {
				const bool del = x < P.xfirst[2];
				scatter_tile<3>(a, del ? a.d_in : a.i_in, del ? a.d_n : a.i_n, x - P.xfirst[del ? 1 : 2], del ? 1 : 2, slot_x, stage_s[wid], runs_s[wid], lane);
			}
		} else if (cur.kind == 2) {
			const int s = source_of(P.yfirst, cur.idx);
			const uint32_t r0 = (cur.idx - P.yfirst[s]) * kPTile;
			const uint32_t valid = min(kPTile, P.ycnt[s] - r0);
			uint2 *out = (uint2 *)(a.peer[s] + a.L.stage) + ((size_t)slot_y * G + a.rank) * a.L.cap_s + r0;     /* the origin's staging area */
			gh::warp_tile_search<kPairs, false, false>(a.table, a.g, y_in(s) + r0, out, valid, true, v, lane, h1, h2);
			pend_y++;
		} else if (cur.kind == 3) {
			gather_tile(a, cur.idx, slot_z, stage_s[wid], runs_s[wid], lane);
		} else if (cur.kind >= 4) {                       /* delete / insert: after every lookup (and delete) of this launch */
			const bool is_delete = cur.kind == 4;
			publish();                                    /* this warp's own finished tiles first: nobody waits on a waiter */
			counter_wait(ws + kWsYDone, P.yfirst[kMaxShards], a, lane);
			if (!is_delete) counter_wait(ws + kWsUDone, P.ufirst[kMaxShards], a, lane);
			const uint32_t *first = is_delete ? P.ufirst : P.vfirst;
			const int s = source_of(first, cur.idx);
			const uint32_t n = is_delete ? P.ucnt[s] : P.vcnt[s], r0 = (cur.idx - first[s]) * kUTile;
			const uint32_t *in = (const uint32_t *)(me + (is_delete ? a.L.inbox_d : a.L.inbox_i)) + 3 * (((size_t)slot_y * G + s) * a.L.cap_u);
			if (kPairs) {                                 /* two lanes per request: 16 requests per round */
#pragma unroll 1
				for (uint32_t r = 0; r < kUTile; r += 16) {
					if (r0 + r >= n) break;               /* warp-uniform */
					const uint32_t i = r0 + r + (lane >> 1);
					const bool have = i < n;
					uint32_t x = 0, y = 0, z = 0;
					if (have) { x = gh::ld_stream_u32(in + 3 * (size_t)i); y = gh::ld_stream_u32(in + 3 * (size_t)i + 1); z = gh::ld_stream_u32(in + 3 * (size_t)i + 2); }
					if (is_delete) {
						const int zc = gh::delete_pair(a.table, a.g, have, x, y, z, lane);
						if (a.st && zc && (lane & 1u) == 0) { atomicAdd(&a.st->del_zeroed, (unsigned long long)zc); atomicAdd(&a.st->del_requests_hit, 1ULL); }
					} else {
						gh::insert_pair(a.table, a.g, have, x, y, z, a.st, lane);
					}
				}
			} else {
#pragma unroll 1
				for (uint32_t r = 0; r < kUTile; r += 32) {
					const uint32_t i = r0 + r + lane;
					if (i < n) {
						const uint32_t x = gh::ld_stream_u32(in + 3 * (size_t)i), y = gh::ld_stream_u32(in + 3 * (size_t)i + 1), z = gh::ld_stream_u32(in + 3 * (size_t)i + 2);
						if (is_delete) {
							const int zc = gh::delete_one<false>(a.table, a.g, x, y, z);
							if (a.st && zc) { atomicAdd(&a.st->del_zeroed, (unsigned long long)zc); atomicAdd(&a.st->del_requests_hit, 1ULL); }
						} else {
							gh::insert_one<false>(a.table, a.g, x, y, z, a.st);
						}
					}
				}
			}
			if (is_delete) pend_u++;
		}
		if (last_of_chunk) {
			publish();
			c0 = n0; c1 = n1;
			n0 = __shfl_sync(0xffffffffu, raw_nn, 0); n1 = n0 + k_nn;
			seen = max(seen, n1);
		}
		t = tn; cur = nxt; v = vn;
	}
	if (in_region) leave_region();
	publish();
	if (a.st) {
		if (h1) atomicAdd(&a.st->search_hits_b1, (unsigned long long)h1);
		if (h2) atomicAdd(&a.st->search_hits_b2, (unsigned long long)h2);
	}
	/* ---- the last CTA leaves the workspace zero and frees the counters of the slot the NEXT launch scatters into */
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(ws + kWsCtas, 1u) == gridDim.x - 1) {
			ws[kWsTicket] = 0; ws[kWsCtas] = 0; ws[kWsRegion] = 0; ws[kWsYDone] = 0; ws[kWsUDone] = 0;
			uint32_t *counts = (uint32_t *)(me + a.L.counts) + ((a.seq + 1u) % 3u) * 24;
			for (int k = 0; k < 24; k++) counts[k] = 0;
			__threadfence();
		}
	}
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

void layout_init(XLayout &L, int G, size_t cap_s, size_t cap_u)
{
	cap_s = align_up(cap_s ? cap_s : 1, kRTile); cap_u = align_up(cap_u ? cap_u : 1, kRTile);
	L.cap_s = (uint32_t)cap_s; L.cap_u = (uint32_t)cap_u;
	size_t o = 0;
	L.inbox_s = o; o += align_up((size_t)3 * G * cap_s * 8, 256);
	L.inbox_d = o; o += align_up((size_t)3 * G * cap_u * 12, 256);
	L.inbox_i = o; o += align_up((size_t)3 * G * cap_u * 12, 256);
	L.stage = o;   o += align_up((size_t)3 * G * cap_s * 8, 256);
	L.cnt = o;     o += 512;
	L.flag = o;    o += 256;
	L.pos = o;     o += align_up((size_t)3 * cap_s, 256);
	L.meta = o;    o += align_up((size_t)3 * (cap_s / kRTile) * 64, 256);
	L.counts = o;  o += 512;
	L.ws = o;      o += 256;
	L.err = o;     o += 256;
	L.total = o;
}

}  // namespace

struct gpuhash_xchg_s {
	gpuhash_geom_t g; void *table; gpuhash_stats_t *stats_d;
	int log2, G, rank; uint32_t hash_mask_total; int shift;
	XLayout L; char *arena; char *peer[kMaxShards]; int have_peers;
	uint32_t seq;
	struct { void *out; uint32_t n; } pend[3];          /* exchange e: pend[e % 3] */
	int grid;
	unsigned long long timeout_ns;
};

extern "C" gpuhash_xchg_t *gpuhash_xchg_create(const gpuhash_geom_t *g, void *table_d, uint32_t hash_mask_total, int log2_shards,
		int my_rank, size_t cap_search, size_t cap_update)
{
	const int G = 1 << log2_shards;
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || !table_d || log2_shards < 0 || log2_shards > 3 || my_rank < 0 || my_rank >= G) return NULL;
	if (cap_search > (1u << 30) || cap_update > (1u << 30)) return NULL;
	int bits = 0; while ((hash_mask_total >> bits) & 1u) bits++;
	if (bits - log2_shards < 0) return NULL;
	gpuhash_xchg_t *x = (gpuhash_xchg_t *)calloc(1, sizeof *x);
	if (!x) return NULL;
	x->g = *g; x->table = table_d; x->log2 = log2_shards; x->G = G; x->rank = my_rank;
	x->hash_mask_total = hash_mask_total; x->shift = bits - log2_shards;
	layout_init(x->L, G, cap_search, cap_update);
	if (cudaMalloc((void **)&x->arena, x->L.total) != cudaSuccess) { free(x); return NULL; }
	/* only the control words need to start at zero; the sync orders the fill before anything a peer or a stream does */
	if (cudaMemset(x->arena + x->L.cnt, 0, x->L.total - x->L.cnt) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
		cudaFree(x->arena); free(x); return NULL;
	}
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const char *e = getenv("GPUHASH_XCHG_CTAS_PER_SM");
	int per_sm = e && atoi(e) > 0 ? atoi(e) : 2;
	x->grid = sms * per_sm;
	x->timeout_ns = 2000000000ULL;
	if (G == 1) { x->peer[0] = x->arena; x->have_peers = 1; }
	return x;
}

extern "C" void *gpuhash_xchg_arena(gpuhash_xchg_t *x, size_t *bytes)
{
	if (!x) return NULL;
	if (bytes) *bytes = x->L.total;
	return x->arena;
}

extern "C" int gpuhash_xchg_set_peers(gpuhash_xchg_t *x, const void *const *peer_arenas)
{
	if (!x || !peer_arenas) return -1;
	for (int r = 0; r < x->G; r++) {
		if (!peer_arenas[r]) return -1;
		x->peer[r] = (char *)peer_arenas[r];
	}
	if (x->peer[x->rank] != x->arena) return -1;
	x->have_peers = 1;
	return 0;
}

extern "C" int gpuhash_xchg_set_stats(gpuhash_xchg_t *x, gpuhash_stats_t *stats_d) { if (!x) return -1; x->stats_d = stats_d; return 0; }

extern "C" int gpuhash_xchg_step(gpuhash_xchg_t *x, const void *search_in, size_t n_search, void *search_out,
		const void *delete_in, size_t n_delete, const void *insert_in, size_t n_insert, void *stream)
{
	if (!x || !x->have_peers) return -1;
	if (n_search > x->L.cap_s || n_delete > x->L.cap_u || n_insert > x->L.cap_u) return -1;
	if ((n_search && (!search_in || !search_out)) || (n_delete && !delete_in) || (n_insert && !insert_in)) return -1;
	if (((uintptr_t)search_in & 7u) || ((uintptr_t)search_out & 7u) || ((uintptr_t)delete_in & 3u) || ((uintptr_t)insert_in & 3u)) return -1;
	const uint32_t seq = x->seq + 1;
	if (seq > 1) {                                       /* every peer's previous launch: its scatter is in my inbox, its results in my staging area */
		int rc = gpuhash_wait_flags((const uint32_t *)(x->arena + x->L.flag), x->G, seq - 1, (uint32_t *)(x->arena + x->L.err), stream);
		if (rc) return rc;
	}
	XArgs a; memset(&a, 0, sizeof a);
	a.table = (gh::Bucket *)x->table;
	a.g.hash_mask = x->g.hash_mask; a.g.block_mask = x->g.block_mask; a.g.algo = x->g.algo; a.g.max_cuckoo = x->g.max_cuckoo; a.g.layout = x->g.layout;
	a.st = (gh::Stats *)x->stats_d;
	for (int r = 0; r < kMaxShards; r++) a.peer[r] = r < x->G ? x->peer[r] : NULL;
	a.L = x->L; a.G = x->G; a.rank = x->rank; a.seq = seq;
	a.hash_mask_total = x->hash_mask_total; a.shift = x->shift;
	a.s_in = (const uint2 *)search_in; a.s_n = (uint32_t)n_search;
	a.d_in = (const uint32_t *)delete_in; a.d_n = (uint32_t)n_delete;
	a.i_in = (const uint32_t *)insert_in; a.i_n = (uint32_t)n_insert;
	a.do_serve = seq >= 2;
	if (seq >= 3) { a.g_out = (uint2 *)x->pend[(seq - 2) % 3].out; a.g_n = x->pend[(seq - 2) % 3].n; }
	a.timeout_ns = x->timeout_ns;
	x->pend[seq % 3].out = search_out; x->pend[seq % 3].n = (uint32_t)n_search;
	x->seq = seq;
	if (x->g.layout == GPUHASH_LAYOUT_PAIRS) xchg_step_kernel<true><<<x->grid, 256, 0, (cudaStream_t)stream>>>(a);
	else                                     xchg_step_kernel<false><<<x->grid, 256, 0, (cudaStream_t)stream>>>(a);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_xchg_flush(gpuhash_xchg_t *x, void *stream)
{
	for (int k = 0; k < 2; k++) {
		int rc = gpuhash_xchg_step(x, NULL, 0, NULL, NULL, 0, NULL, 0, stream);
		if (rc) return rc;
	}
	return 0;
}

extern "C" unsigned gpuhash_xchg_seq(const gpuhash_xchg_t *x) { return x ? x->seq : 0u; }

extern "C" int gpuhash_xchg_error(gpuhash_xchg_t *x)
{
	if (!x) return -1;
	uint32_t e = 0;
	cudaError_t rc = cudaMemcpy(&e, x->arena + x->L.err, sizeof e, cudaMemcpyDeviceToHost);
	if (rc != cudaSuccess) return (int)rc;
	return e ? -3 : 0;
}

extern "C" void gpuhash_xchg_destroy(gpuhash_xchg_t *x)
{
	if (!x) return;
	cudaFree(x->arena);
	free(x);
}
