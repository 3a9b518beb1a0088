/*
 * gpuhash_xchg.cu -- the sharded index as ONE kernel per scheduler cycle: routing, lookup and un-routing fused,
 * warp-specialised and software-pipelined over NVLink peer memory (BASELINE.json north_star (d)).
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
 * is the whole inter-GPU synchronisation.
 *
 * Inside the launch the three kinds of work are independent, and they want different things from the SM.  A lookup is
 * bound by random line fills and needs many loads in flight: 16 warps with four 32 B table loads per lane, ~100
 * registers each (the cycle kernel's shape, gpuhash_kernels.cuh).  Routing is streaming traffic plus shuffling through
 * shared memory: little state, long dependent chains.  Run by the same warps one after the other, every microsecond a
 * warp routes is a microsecond it has no probes in flight (first version of this file: 12.5 Gops/s on one GPU against
 * 20.4 for the lookups alone).  So the CTA (one per SM) is WARP-SPECIALISED: 16 lookup warps + 8 (or 4) router warps,
 * and the register file is re-cut between them at kernel entry with setmaxnreg (the routers give registers up, the
 * lookup warps take them), so that the routers are extra residents, not a tax on the lookups.  Each role has its own
 * ticket counter:
 *      lookup warps   lookup tiles (64 requests), then delete tiles, then insert tiles (64 requests; they wait --
 *                     device-scope counters -- for every lookup / delete of this launch, as a worker's insert follows
 *                     its search in the reference)
 *      router warps   scatter tiles and gather tiles (256 requests) alternating; the raw requests of the NEXT scatter
 *                     tile are already on their way into shared memory (cp.async.bulk + mbarrier, two stages)
 * The routing traffic (NVLink stores, coalesced reads) thus rides in the shadow of the lookups' line fills.
 *
 * Buffers are triple-buffered by exchange number (slot = e mod 3): a peer may already be scattering exchange j+1
 * into slot (j+1) mod 3 of my inbox while I still serve exchange j-1 from slot (j-1) mod 3; slot e mod 3 is written
 * again by launch e+3 at the earliest, which waits for the flags of launch e+2, raised after everyone's use of e.
 *
 * A scatter tile (one warp): 256 requests, owner = top bits of bucket 1 (== of bucket 2 and of every eviction target,
 * gpu_hash.h:67-69), rank inside (tile, owner) by ballots, one global atomic per (tile, owner) reserves the run in the
 * owner's region, the tile is sorted through shared memory and every run leaves as contiguous stores.  The map the gather
 * needs: one 32-bit word per request (owner << 28 | index in the owner's region = where its result will be staged).  A gather
 * tile issues the eight loads of each lane at once (the 256 results of a tile lie in at most G contiguous runs, 2 KB in all)
 * and writes the results in request order, 512 B per warp store.
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
constexpr uint32_t kMaxClaim = 4;          /* lookup-side tiles per ticket at most */

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
	size_t where;          /* [3][cap_s] u32             owner << 28 | index in the owner's region: where request i's result will be */
	size_t counts;         /* [3][3][8] u32              slots handed out per (kind, owner); zero when the slot is free */
	size_t ws;             /* u32 words, see kWs*        tickets and counters of one launch; zero between launches */
	size_t err;            /* u32                        sticky: a wait inside a kernel timed out */
	size_t total;
	uint32_t cap_s, cap_u;
};

/* workspace words, one 32 B sector each */
enum { kWsTicketL = 0, kWsTicketR = 8, kWsCtas = 16, kWsRegion = 24, kWsYDone = 32, kWsUDone = 40, kWsWords = 48 };

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

struct Plan {                              /* built once per CTA in shared memory */
	uint32_t xfirst[4];                    /* scatter tiles: search | delete | insert */
	uint32_t ycnt[kMaxShards], yfirst[kMaxShards + 1];      /* lookup tiles per source */
	uint32_t ucnt[kMaxShards], ufirst[kMaxShards + 1];      /* delete tiles per source */
	uint32_t vcnt[kMaxShards], vfirst[kMaxShards + 1];      /* insert tiles per source */
	uint32_t nZ;
};

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

/* Leaving the part of a launch that peers depend on (scatter, lookups, gather): everything this warp stored for peers
 * is complete at system scope before it counts itself; the last warp of the grid to leave hands the peers my counts
 * and raises my flag. */
__device__ __forceinline__ void leave_region(const XArgs &a, uint32_t nwarps, unsigned lane)
{
	char *me = a.peer[a.rank];
	uint32_t *ws = (uint32_t *)(me + a.L.ws);
	const uint32_t slot_x = a.seq % 3u;
	__syncwarp();
	uint32_t last = 0;
	if (lane == 0) {
		__threadfence_system();
		last = atomicAdd(ws + kWsRegion, 1u) == nwarps - 1 ? 1u : 0u;
	}
	last = __shfl_sync(0xffffffffu, last, 0);
	if (last && lane < (unsigned)a.G) {
		__threadfence();
		const uint32_t *counts = (const uint32_t *)(me + a.L.counts) + slot_x * 24;
		volatile uint32_t *pc = (volatile uint32_t *)(a.peer[lane] + a.L.cnt) + slot_x * 24;
		pc[a.rank] = ld_relaxed_gpu(counts + lane);
		pc[8 + a.rank] = ld_relaxed_gpu(counts + 8 + lane);
		pc[16 + a.rank] = ld_relaxed_gpu(counts + 16 + lane);
		__threadfence_system();
		asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)(a.peer[lane] + a.L.flag) + a.rank), "r"(a.seq) : "memory");
	}
}

/* ================================================ router warps ================================================ */

struct RouterSmem {                        /* per router warp */
	uint32_t raw[2][kRTile * 3];           /* the next scatter tiles as they lie in the request array (bulk copies land here) */
	uint32_t sorted[kRTile * 3];           /* a tile sorted by owner (scatter) / the result runs of a tile (gather) */
	uint32_t runs[48];                     /* [0..8] run starts inside the sorted tile, [16..31] per owner: pointer to where sorted position 0 would go, [32..39] its index in the owner's region */
	unsigned long long bar[2];
};

/* what a router ticket means: even -> scatter tile t/2, odd -> gather tile t/2 (either may not exist) */
enum { kNone = 0, kScatterSearch, kScatterDelete, kScatterInsert, kGather };
struct RTile { int kind; uint32_t idx; const uint32_t *in; uint32_t n; int words; };   /* idx: tile inside its request array / result array */

template <int kWords>
__device__ __forceinline__ void scatter_tile(const XArgs &a, const uint32_t *raw, uint32_t tile, uint32_t tile_n, int kind, uint32_t slot,
		uint32_t *sorted, uint32_t *runs, unsigned lane)
{
	const uint32_t t0 = tile * kRTile;
	const int G = a.G;
	/* rank inside (tile, owner): one shared-memory atomic per request on the owner's counter (any order inside an owner will do:
	 * where[] records the place).  Request k = 32 r + lane.  (First version: 8 ballots per round and owner with warp-uniform
	 * running counts -- 450 of the tile's ~1000 instructions; a router warp is bound by its own instruction latency.) */
	uint32_t key[8];                       /* owner << 16 | rank */
	if (lane < kMaxShards) runs[40 + lane] = 0;
	__syncwarp();
#pragma unroll
	for (int r = 0; r < 8; r++) {
		const uint32_t k = 32 * r + lane;
		uint32_t d = 0xffu, rk = 0;
		if (k < tile_n) {
			d = (raw[kWords * k + 1] & a.hash_mask_total) >> a.shift;              /* owner = top bits of bucket 1 */
			rk = G == 1 ? k : atomicAdd(&runs[40 + d], 1u);
		}
		key[r] = (d << 16) | rk;
	}
	__syncwarp();
	/* lane o: reserve owner o's run in its region; the atomic's round trip hides behind the sorting below */
	uint32_t mine = 0;
	if (lane < kMaxShards) mine = G == 1 ? (lane == 0 ? tile_n : 0u) : runs[40 + lane];
	uint32_t inc = mine;
#pragma unroll
	for (int dd = 1; dd < kMaxShards; dd <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, dd); if ((int)lane >= dd) inc += o; }
	const uint32_t off_mine = inc - mine;
	uint32_t base = 0;
	uint32_t *counts = (uint32_t *)(a.peer[a.rank] + a.L.counts) + (slot * 3 + kind) * 8;
	if (lane < (unsigned)G && mine) base = atomicAdd(counts + lane, mine);
	if (lane <= kMaxShards) runs[lane] = lane < kMaxShards ? off_mine : tile_n;      /* off[0..8] */
	__syncwarp();
#pragma unroll
	for (int r = 0; r < 8; r++) {
		const uint32_t k = 32 * r + lane;
		if (k < tile_n) {
			const uint32_t p = runs[key[r] >> 16] + (key[r] & 0xffffu);
#pragma unroll
			for (int j = 0; j < kWords; j++) sorted[p * kWords + j] = raw[kWords * k + j];
			key[r] = (key[r] & 0xffff0000u) | p;           /* owner << 16 | place in the sorted tile */
		}
	}
	if (lane < kMaxShards) runs[32 + lane] = base - off_mine;      /* index in the owner's region = this + place (mod 2^32) */
	/* lane o: where sorted position q of owner o goes = dst[o] + q requests (the run start is folded into the pointer) */
	if (lane < kMaxShards) {
		const size_t region = ((size_t)slot * G + a.rank) * (kWords == 2 ? a.L.cap_s : a.L.cap_u);
		const size_t box = kind == 0 ? a.L.inbox_s : (kind == 1 ? a.L.inbox_d : a.L.inbox_i);
		char *dst = lane < (unsigned)G ? a.peer[lane] + box + (region + base) * (size_t)(4 * kWords) - (size_t)off_mine * (4 * kWords) : nullptr;
		*(char **)(runs + 16 + 2 * lane) = dst;
	}
	__syncwarp();
	if (kind == 0) {                    /* the gather's map: one coalesced 128 B store per round */
		uint32_t *where = (uint32_t *)(a.peer[a.rank] + a.L.where) + (size_t)slot * a.L.cap_s + t0;
#pragma unroll
		for (int r = 0; r < 8; r++) {
			const uint32_t k = 32 * r + lane;
			if (k < tile_n) { const uint32_t o = key[r] >> 16; where[k] = (o << 28) | (runs[32 + o] + (key[r] & 0xffffu)); }
		}
	}
	/* runs out: neighbouring lanes share a run -> contiguous stores, local or over NVLink */
	const uint32_t t1 = runs[1], t2 = runs[2], t3 = runs[3], t4 = runs[4], t5 = runs[5], t6 = runs[6], t7 = runs[7];   /* run starts */
	auto owner_at = [&](uint32_t q) -> int {
		return (int)(q >= t1) + (int)(q >= t2) + (int)(q >= t3) + (int)(q >= t4) + (int)(q >= t5) + (int)(q >= t6) + (int)(q >= t7);
	};
	if (kWords == 2) {
#pragma unroll
		for (int it = 0; it < 8; it++) {
			const uint32_t q = 32 * it + lane;
			if (q < tile_n) {
				const int o = owner_at(q);
				uint2 *dst = *(uint2 **)(runs + 16 + 2 * o) + q;
				*dst = *(const uint2 *)(sorted + 2 * q);
			}
		}
	} else {
#pragma unroll 8
		for (int it = 0; it < 24; it++) {
			const uint32_t wq = 32 * it + lane;
			if (wq < 3 * tile_n) {
				const uint32_t q = wq / 3;
				const int o = owner_at(q);
				uint32_t *dst = *(uint32_t **)(runs + 16 + 2 * o) + wq;
				*dst = sorted[wq];
			}
		}
	}
	__syncwarp();                                          /* sorted / runs are free again */
}

/* ---- gather: one warp, one tile of 256 results back into request order ----
 * where[i] says which staging region (owner) and which index holds request i's result.  The 256 results of a tile lie in at
 * most G contiguous runs (2 KB in all), so the eight 8-byte loads of a lane -- all in flight at once, L1-allocating so that the
 * lanes' neighbours in a line share its fill -- move each line once; the stores are 512 B per warp in request order.  No shared
 * memory, no dependence on anything but where[], which is loaded one tile ahead. */
struct GPre { uint2 w[4]; };               /* where[] of this lane's requests 64 it + 2 lane, + 1 */

__device__ __forceinline__ GPre gather_prefetch(const XArgs &a, uint32_t tile, uint32_t slot, unsigned lane)
{
	const uint32_t t0 = tile * kRTile, tile_n = min(kRTile, a.g_n - t0);
	const uint32_t *where = (const uint32_t *)(a.peer[a.rank] + a.L.where) + (size_t)slot * a.L.cap_s + t0;
	GPre g;
#pragma unroll
	for (int it = 0; it < 4; it++) {
		const uint32_t k = 64 * it + 2 * lane;
		g.w[it] = make_uint2(0u, 0u);
		if (k + 1 < tile_n) g.w[it] = gh::ld_stream_u2((const uint2 *)(where + k));
		else if (k < tile_n) g.w[it].x = gh::ld_stream_u32(where + k);
	}
	return g;
}

__device__ __forceinline__ void gather_tile(const XArgs &a, uint32_t tile, uint32_t slot, const GPre &pre, unsigned lane)
{
	const uint32_t t0 = tile * kRTile, tile_n = min(kRTile, a.g_n - t0);
	const uint2 *mine = (const uint2 *)(a.peer[a.rank] + a.L.stage) + (size_t)slot * a.G * a.L.cap_s;
	uint2 v[8];
#pragma unroll
	for (int it = 0; it < 4; it++) {
		const uint32_t k = 64 * it + 2 * lane;
		v[2 * it] = make_uint2(0u, 0u); v[2 * it + 1] = make_uint2(0u, 0u);
		if (k < tile_n) {
			const uint2 *src = mine + (size_t)(pre.w[it].x >> 28) * a.L.cap_s + (pre.w[it].x & 0x0fffffffu);
			asm volatile("ld.global.ca.v2.u32 {%0,%1}, [%2];" : "=r"(v[2 * it].x), "=r"(v[2 * it].y) : "l"(src));
		}
		if (k + 1 < tile_n) {
			const uint2 *src = mine + (size_t)(pre.w[it].y >> 28) * a.L.cap_s + (pre.w[it].y & 0x0fffffffu);
			asm volatile("ld.global.ca.v2.u32 {%0,%1}, [%2];" : "=r"(v[2 * it + 1].x), "=r"(v[2 * it + 1].y) : "l"(src));
		}
	}
	uint2 *out = a.g_out + t0;
	const bool vec = ((uintptr_t)a.g_out & 15u) == 0;
#pragma unroll
	for (int it = 0; it < 4; it++) {
		const uint32_t k = 64 * it + 2 * lane;
		if (k + 1 < tile_n) {
			if (vec) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out + k), "r"(v[2 * it].x), "r"(v[2 * it].y), "r"(v[2 * it + 1].x), "r"(v[2 * it + 1].y) : "memory");
			else { gh::st_stream_u2(out + k, v[2 * it]); gh::st_stream_u2(out + k + 1, v[2 * it + 1]); }
		} else if (k < tile_n) {
			gh::st_stream_u2(out + k, v[2 * it]);
		}
	}
}

__device__ __forceinline__ void router_loop(const XArgs &a, const Plan &P, RouterSmem &S, uint32_t nwarps, unsigned lane)
{
	char *me = a.peer[a.rank];
	uint32_t *ws = (uint32_t *)(me + a.L.ws);
	const uint32_t slot_x = a.seq % 3u, slot_z = (a.seq + 1u) % 3u;
	const uint32_t nX = P.xfirst[3], total = 2 * max(nX, P.nZ);
	if (lane == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gh::smem_u32(&S.bar[0])));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gh::smem_u32(&S.bar[1])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();
	auto decode = [&](uint32_t t) -> RTile {
		RTile r; r.kind = kNone; r.idx = t >> 1; r.in = nullptr; r.n = 0; r.words = 2;
		if (t >= total) return r;
		if (t & 1u) { if (r.idx < P.nZ) r.kind = kGather; return r; }
		if (r.idx >= nX) return r;
		if (r.idx < P.xfirst[1])      { r.kind = kScatterSearch; r.in = (const uint32_t *)a.s_in; r.n = a.s_n; }
		else if (r.idx < P.xfirst[2]) { r.kind = kScatterDelete; r.in = a.d_in; r.n = a.d_n; r.idx -= P.xfirst[1]; r.words = 3; }
		else                          { r.kind = kScatterInsert; r.in = a.i_in; r.n = a.i_n; r.idx -= P.xfirst[2]; r.words = 3; }
		return r;
	};
	/* bring a scatter tile's raw requests into stage s: one bulk copy when the piece is 16 B-granular, plain loads else.
	 * Returns 1 if the mbarrier of the stage will complete for it. */
	auto fetch = [&](const RTile &tl, int s) -> uint32_t {
		if (tl.kind == kNone || tl.kind == kGather) return 0u;
		const uint32_t t0 = tl.idx * kRTile, tile_n = min(kRTile, tl.n - t0);
		const uint32_t bytes = tile_n * tl.words * 4;
		const uint32_t *src = tl.in + (size_t)tl.words * t0;
		if ((((uintptr_t)src | bytes) & 15u) == 0) {
			if (lane == 0) {
				const uint32_t b = gh::smem_u32(&S.bar[s]);
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(bytes) : "memory");
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					:: "r"(gh::smem_u32(&S.raw[s][0])), "l"(src), "r"(bytes), "r"(b) : "memory");
			}
			return 1u;
		}
		for (uint32_t q = lane; q < tile_n * tl.words; q += 32) S.raw[s][q] = gh::ld_stream_u32(src + q);
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      /* a later bulk copy into this stage follows generic writes */
		return 0u;
	};
	uint32_t raw_t = lane == 0 ? atomicAdd(ws + kWsTicketR, 1u) : 0u;
	uint32_t raw_n = lane == 0 ? atomicAdd(ws + kWsTicketR, 1u) : 0u;
	uint32_t t = __shfl_sync(0xffffffffu, raw_t, 0);
	RTile cur = decode(t);
	int s = 0;
	uint32_t phases = 0;
	uint32_t bulk = fetch(cur, 0);
	GPre gp = {};
	if (cur.kind == kGather) gp = gather_prefetch(a, cur.idx, slot_z, lane);
	while (t < total) {
		const uint32_t tn = __shfl_sync(0xffffffffu, raw_n, 0);
		raw_n = lane == 0 ? atomicAdd(ws + kWsTicketR, 1u) : 0u;      /* the ticket after next: its latency is nobody's critical path */
		const RTile nxt = decode(tn);
		__syncwarp();                                      /* stage s^1 was read (sorted out of it) two tiles ago by every lane */
		const uint32_t bulk_n = fetch(nxt, s ^ 1);
		GPre gp_n = {};
		if (nxt.kind == kGather) gp_n = gather_prefetch(a, nxt.idx, slot_z, lane);
		if (cur.kind == kGather) {
			gather_tile(a, cur.idx, slot_z, gp, lane);
		} else if (cur.kind) {
			if (bulk) { gh::mbar_wait(gh::smem_u32(&S.bar[s]), (phases >> s) & 1u); phases ^= 1u << s; }
			else __syncwarp();
			const uint32_t tile_n = min(kRTile, cur.n - cur.idx * kRTile);
			/* scatter_tile's `kind`: 0 search, 1 delete, 2 insert (index of the inbox and of the counters) */
			if (cur.kind == kScatterSearch) scatter_tile<2>(a, S.raw[s], cur.idx, tile_n, 0, slot_x, S.sorted, S.runs, lane);
			else                            scatter_tile<3>(a, S.raw[s], cur.idx, tile_n, cur.kind == kScatterDelete ? 1 : 2, slot_x, S.sorted, S.runs, lane);
		}
		t = tn; cur = nxt; bulk = bulk_n; gp = gp_n; s ^= 1;
	}
	leave_region(a, nwarps, lane);
}

/* ================================================ lookup warps ================================================ */

template <bool kPairs>
__device__ __forceinline__ void lookup_loop(const XArgs &a, const Plan &P, uint32_t nwarps, uint32_t nl /* lookup warps of the grid */, unsigned lane)
{
	char *me = a.peer[a.rank];
	uint32_t *ws = (uint32_t *)(me + a.L.ws);
	const int G = a.G;
	const uint32_t slot_y = (a.seq + 2u) % 3u;
	const uint32_t nY = P.yfirst[kMaxShards], nU = P.ufirst[kMaxShards], nV = P.vfirst[kMaxShards];
	const uint32_t total = nY + nU + nV;
	uint32_t h1 = 0, h2 = 0;
	/* tiles this warp finished and has not published yet.  Published (fence + one add) when the warp leaves a PHASE -- lookups,
	 * deletes -- not per claim: the fence has to wait for the warp's result stores, and per claim it cost the lookup warps 12 % of
	 * their time (ncu, stall_membar).  A delete / insert tile therefore waits until every lookup warp of the grid (one CTA per
	 * SM: all resident) has run out of lookup tiles, which costs one tile time at the phase boundary. */
	uint32_t pend_y = 0, pend_u = 0;
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
	auto claim_size = [&](uint32_t seen) -> uint32_t {
		const uint32_t rem = total > seen ? total - seen : 0u;
		return min(kMaxClaim, max(1u, rem / (6u * nl)));
	};
	auto claim_issue = [&](uint32_t k) -> uint32_t { return lane == 0 ? atomicAdd(ws + kWsTicketL, k) : 0u; };
	auto y_in = [&](int s) -> const uint2 * { return (const uint2 *)(me + a.L.inbox_s) + ((size_t)slot_y * G + s) * a.L.cap_s; };
	/* the requests of a lookup tile (this lane's 16 B): in flight while the tile before it is worked on */
	auto prefetch = [&](uint32_t t) -> uint4 {
		if (t >= nY) return make_uint4(0u, 0u, 0u, 0u);
		const int s = source_of(P.yfirst, t);
		const uint32_t r0 = (t - P.yfirst[s]) * kPTile;
		return gh::warp_tile_load<false>(y_in(s) + r0, min(kPTile, P.ycnt[s] - r0), lane);
	};

	uint32_t k_cur = claim_size(0), k_nxt = k_cur;
	const uint32_t raw_a = claim_issue(k_cur), raw_b = claim_issue(k_nxt);
	uint32_t c0 = __shfl_sync(0xffffffffu, raw_a, 0), c1 = c0 + k_cur;
	uint32_t n0 = __shfl_sync(0xffffffffu, raw_b, 0), n1 = n0 + k_nxt;
	uint32_t seen = n1, raw_nn = 0, k_nn = 1;
	uint32_t t = c0;
	uint4 v = prefetch(t);
	while (t < total) {
		if (t == c0) { k_nn = claim_size(seen); raw_nn = claim_issue(k_nn); }
		if (in_region && t >= nY) { publish(); leave_region(a, nwarps, lane); in_region = false; }
		const bool last_of_chunk = t + 1 >= c1;
		const uint32_t tn = last_of_chunk ? n0 : t + 1;
		const uint4 vn = prefetch(tn);
		if (t < nY) {
			const int s = source_of(P.yfirst, t);
			const uint32_t r0 = (t - P.yfirst[s]) * kPTile;
			const uint32_t valid = min(kPTile, P.ycnt[s] - r0);
			uint2 *out = (uint2 *)(a.peer[s] + a.L.stage) + ((size_t)slot_y * G + a.rank) * a.L.cap_s + r0;     /* the origin's staging area */
			gh::warp_tile_search<kPairs, false, false>(a.table, a.g, y_in(s) + r0, out, valid, true, v, lane, h1, h2);
			pend_y++;
		} else {                                          /* delete / insert: after every lookup (and delete) of this launch */
			const bool is_delete = t < nY + nU;
			publish();                                    /* this warp's own finished tiles first: nobody waits on a waiter */
			counter_wait(ws + kWsYDone, nY, a, lane);
			if (!is_delete) counter_wait(ws + kWsUDone, nU, a, lane);
			const uint32_t *first = is_delete ? P.ufirst : P.vfirst;
			const uint32_t idx = is_delete ? t - nY : t - nY - nU;
			const int s = source_of(first, idx);
			const uint32_t n = is_delete ? P.ucnt[s] : P.vcnt[s], r0 = (idx - first[s]) * kUTile;
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
			c0 = n0; c1 = n1;
			n0 = __shfl_sync(0xffffffffu, raw_nn, 0); n1 = n0 + k_nn;
			seen = max(seen, n1);
		}
		t = tn; v = vn;
	}
	publish();
	if (in_region) leave_region(a, nwarps, lane);
	if (a.st) {
		if (h1) atomicAdd(&a.st->search_hits_b1, (unsigned long long)h1);
		if (h2) atomicAdd(&a.st->search_hits_b2, (unsigned long long)h2);
	}
}

/* CTA shapes: kLW lookup warps + kRW router warps, one CTA per SM.  The kernel is compiled for 65536 / threads registers per
 * thread (rounded down to 8); where the two roles want different budgets the register file is re-cut at kernel entry with
 * setmaxnreg (warpgroup-wide, multiples of 8): the routers drop to kRReg, which lets every lookup warp rise to kLReg
 * (0 = leave the launch allocation alone).
 *      16 + 4   640 threads, 96 regs at launch -> routers 64, lookups 104     (128*64 + 512*104 = 61440 = 640*96)
 *      16 + 8   768 threads, 80                -> routers 56, lookups 88      (256*56 + 512*88  = 59392 <= 61440)
 *      24 + 8  1024 threads, 64                -> as launched: more lookup warps with the search kernel's 64 registers */
template <bool kPairs, int kLW, int kRW, int kLReg, int kRReg>
__global__ void __launch_bounds__(32 * (kLW + kRW), 1)
xchg_step_kernel(XArgs a)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	__shared__ Plan P;
	const unsigned lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
	const int G = a.G;
	char *me = a.peer[a.rank];
	uint32_t *ws = (uint32_t *)(me + a.L.ws);

	if (threadIdx.x == 0) {
		P.xfirst[0] = 0;
		P.xfirst[1] = (a.s_n + kRTile - 1) / kRTile;
		P.xfirst[2] = P.xfirst[1] + (a.d_n + kRTile - 1) / kRTile;
		P.xfirst[3] = P.xfirst[2] + (a.i_n + kRTile - 1) / kRTile;
		const uint32_t *cnt = (const uint32_t *)(me + a.L.cnt) + ((a.seq + 2u) % 3u) * 24;     /* written by the peers one launch ago */
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
	}
	__syncthreads();
	const uint32_t nwarps = gridDim.x * (kLW + kRW);
	if (wid >= (unsigned)kLW) {
		if (kRReg) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRReg));
		RouterSmem *S = (RouterSmem *)smem_raw + (wid - kLW);
		router_loop(a, P, *S, nwarps, lane);
	} else {
		if (kLReg) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kLReg));
		lookup_loop<kPairs>(a, P, nwarps, gridDim.x * kLW, lane);
	}
	/* ---- the last CTA leaves the workspace zero and frees the counters of the slot the NEXT launch scatters into */
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(ws + kWsCtas, 1u) == gridDim.x - 1) {
			ws[kWsTicketL] = 0; ws[kWsTicketR] = 0; ws[kWsCtas] = 0; ws[kWsRegion] = 0; ws[kWsYDone] = 0; ws[kWsUDone] = 0;
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
	L.where = o;   o += align_up((size_t)3 * cap_s * 4, 256);
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
	int grid, shape;
	unsigned long long timeout_ns;
};

extern "C" gpuhash_xchg_t *gpuhash_xchg_create(const gpuhash_geom_t *g, void *table_d, uint32_t hash_mask_total, int log2_shards,
		int my_rank, size_t cap_search, size_t cap_update)
{
	const int G = 1 << log2_shards;
	if (!g || g->layout > GPUHASH_LAYOUT_REFERENCE || !table_d || log2_shards < 0 || log2_shards > 3 || my_rank < 0 || my_rank >= G) return NULL;
	if (cap_search > (1u << 28) - kRTile || cap_update > (1u << 28) - kRTile) return NULL;      /* where[] keeps 28 bits of index */
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
	const char *e = getenv("GPUHASH_XCHG_SHAPE");        /* "16x8" (default: 297 vs 315 us per cycle on 2 GPUs), "16x4", "24x8" */
	x->shape = e && !strcmp(e, "16x4") ? 0 : (e && !strcmp(e, "24x8") ? 2 : 1);
	x->grid = sms;                                       /* one warp-specialised CTA per SM */
	const int smem8 = 8 * (int)sizeof(RouterSmem);       /* > 48 KB: opt in */
	if (cudaFuncSetAttribute(xchg_step_kernel<true, 16, 8, 88, 56>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8) != cudaSuccess
			|| cudaFuncSetAttribute(xchg_step_kernel<false, 16, 8, 88, 56>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8) != cudaSuccess
			|| cudaFuncSetAttribute(xchg_step_kernel<true, 24, 8, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8) != cudaSuccess
			|| cudaFuncSetAttribute(xchg_step_kernel<false, 24, 8, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8) != cudaSuccess) {
		cudaFree(x->arena); free(x); return NULL;
	}
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
	const bool pairs = x->g.layout == GPUHASH_LAYOUT_PAIRS;
	cudaStream_t s = (cudaStream_t)stream;
	const size_t sm8 = 8 * sizeof(RouterSmem), sm4 = 4 * sizeof(RouterSmem);
	if (x->shape == 1) {
		if (pairs) xchg_step_kernel<true, 16, 8, 88, 56><<<x->grid, 32 * 24, sm8, s>>>(a);
		else       xchg_step_kernel<false, 16, 8, 88, 56><<<x->grid, 32 * 24, sm8, s>>>(a);
	} else if (x->shape == 2) {
		if (pairs) xchg_step_kernel<true, 24, 8, 0, 0><<<x->grid, 32 * 32, sm8, s>>>(a);
		else       xchg_step_kernel<false, 24, 8, 0, 0><<<x->grid, 32 * 32, sm8, s>>>(a);
	} else {
		if (pairs) xchg_step_kernel<true, 16, 4, 104, 64><<<x->grid, 32 * 20, sm4, s>>>(a);
		else       xchg_step_kernel<false, 16, 4, 104, 64><<<x->grid, 32 * 20, sm4, s>>>(a);
	}
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
