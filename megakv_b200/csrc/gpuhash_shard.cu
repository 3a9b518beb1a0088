/*
 * gpuhash_shard.cu -- device side of the sharded index (BASELINE.json north_star (d)).
 *
 * The reference is single-GPU (assert(config->scheduler_num == 1), src/mega.c:410; no cudaSetDevice
 * anywhere).  What makes sharding exact is a property of its hash functions: both candidate buckets of
 * a key, and every bucket an eviction chain can reach, share the top IBLOCK_P = 3 bits of the bucket
 * index (BLOCK_HASH_MASK keeps them, gpu_hash.h:67-69, gpu_hash.cu:66-67,334-335).  So G in {2,4,8}
 * contiguous bucket ranges are closed under search / insert / delete, and shard g can hold its range as
 * a local table (gpuhash_geom_init_shard) whose results equal the single-table oracle's.
 *
 * Per batch, per rank:   scatter requests by owner  ->  exchange  ->  local kernel on what arrived
 *                        (searches only:)  exchange results back  ->  gather into request order
 * The exchange is either NCCL (torch.distributed all_to_all, megakv_b200/sharded.py) or -- the fused
 * path -- the scatter kernel storing straight into the owners' inboxes over NVLink and the lookup kernel
 * storing results straight into the origin's staging area (peer pointers from CUDA IPC), with
 * sequence-numbered flags instead of host synchronisation.
 *
 * Regions have a fixed capacity `cap` (requests per source per batch), so no prefix sum over the whole
 * batch is needed: region d of a rank starts at d * cap and a slot is claimed with one shared-memory
 * atomic per request + one global atomic per (CTA, destination).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <cuda.h>             /* types of the stream memory operations only; resolved at run time, no -lcuda */
#include <cuda_runtime.h>
#include "gpuhash_ex.h"
#include "gpuhash_kernels.cuh"

namespace {

constexpr int kMaxShards = 8;

struct Ptrs { void *p[kMaxShards]; };

/* data another GPU stored into this GPU's memory (or that lives in a peer's): never from L1 */
__device__ __forceinline__ uint2 ld_u2_sys(const void *p)
{
	uint2 v;
	asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ uint64_t globaltimer_ns()
{
	uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t;
}

/* Wait until *flag >= want (flags only grow: they carry the batch sequence number).  System-scope
 * acquire load: the flag is written by another GPU.  Gives up after `timeout_ns` and reports through
 * *err so that a dead peer turns into an error code, not a hung box. */
__device__ __forceinline__ bool wait_flag(const volatile uint32_t *flag, uint32_t want, uint64_t timeout_ns, uint32_t *err)
{
	/* poll with relaxed loads (an acquire at system scope invalidates this SM's L1 every time it is issued, which
	 * the kernels running next to the waiter pay for), then ONE acquire fence once the flag is there */
	uint64_t t0 = globaltimer_ns();
	for (;;) {
		uint32_t v;
		asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
		if ((int32_t)(v - want) >= 0) break;
		if (globaltimer_ns() - t0 > timeout_ns) { atomicExch(err, 1u); return false; }
		__nanosleep(256);
	}
	asm volatile("fence.acq_rel.sys;" ::: "memory");
	return true;
}

/* ---- K1: scatter a batch into per-owner regions ------------------------------------------------- */
template <int kWords>
__global__ void __launch_bounds__(256)
route_scatter_kernel(const uint32_t *__restrict__ in, size_t n, uint32_t hash_mask_total, int shift, int G,
		Ptrs dst, uint32_t *counts /* [G], zero on entry */, uint32_t *perm /* [G][cap] or NULL */, size_t cap)
{
	__shared__ uint32_t blk_count[kMaxShards], blk_base[kMaxShards];
	for (size_t tile = (size_t)blockIdx.x * blockDim.x; tile < n; tile += (size_t)gridDim.x * blockDim.x) {
		if (threadIdx.x < kMaxShards) blk_count[threadIdx.x] = 0;
		__syncthreads();
		const size_t i = tile + threadIdx.x;
		uint32_t w[kWords]; uint32_t d = 0, rank = 0;
		if (i < n) {
#pragma unroll
			for (int k = 0; k < kWords; k++) w[k] = gh::ld_stream_u32(in + kWords * i + k);
			d = (w[1] & hash_mask_total) >> shift;                   /* owner = top bits of bucket 1 (== of bucket 2) */
			rank = atomicAdd(&blk_count[d], 1u);
		}
		__syncthreads();
		if (threadIdx.x < G && blk_count[threadIdx.x])
			blk_base[threadIdx.x] = atomicAdd(&counts[threadIdx.x], blk_count[threadIdx.x]);
		__syncthreads();
		if (i < n) {
			const size_t slot = (size_t)blk_base[d] + rank;          /* < cap as long as cap >= n */
			uint32_t *q = (uint32_t *)dst.p[d] + kWords * slot;
#pragma unroll
			for (int k = 0; k < kWords; k++) q[k] = w[k];            /* plain stores: local HBM or a peer over NVLink */
			if (perm) perm[(size_t)d * cap + slot] = (uint32_t)i;
		}
		__syncthreads();
	}
}

/* Publish this rank's per-owner counts to the owners: inbox_count[src=my_rank] on peer d, then the sequence
 * flag.  One thread per destination; the fence orders the scatter kernel's stores (previous kernel on the
 * same stream, already complete) and the count before the flag at system scope. */
__global__ void route_publish_kernel(const uint32_t *counts, int G, int my_rank, Ptrs peer_count /* [G] -> uint32[G] */,
		Ptrs peer_flag /* [G] -> uint32[G] */, uint32_t seq)
{
	int d = threadIdx.x;
	if (d >= G) return;
	((volatile uint32_t *)peer_count.p[d])[my_rank] = counts[d];
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)peer_flag.p[d] + my_rank), "r"(seq) : "memory");
}

/* ---- K2: look up everything that arrived, region by region -------------------------------------- */
/* seg_count may be written by peers (fused path): then wait_seq != 0 and thread 0 of every CTA first waits
 * for all G source flags to reach wait_seq. */
__global__ void __launch_bounds__(256)
search_segments_kernel(const gh::Bucket *__restrict__ table, gh::Geom g, int G, Ptrs seg_in, const uint32_t *seg_count,
		Ptrs seg_out, const uint32_t *flags, uint32_t wait_seq, uint32_t *err)
{
	__shared__ uint32_t prefix[kMaxShards + 1];
	__shared__ int ok;
	if (threadIdx.x == 0) {
		ok = 1;
		if (wait_seq) for (int s = 0; s < G && ok; s++) ok = wait_flag(flags + s, wait_seq, 2000000000ULL, err);
		uint32_t acc = 0;
		for (int s = 0; s < G; s++) { prefix[s] = acc; acc += ((const volatile uint32_t *)seg_count)[s]; }
		prefix[G] = acc;
	}
	__syncthreads();
	if (!ok) return;
	const uint32_t total = prefix[G];
	int s = 0;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
		while (e >= prefix[s + 1]) s++;
		const uint32_t j = e - prefix[s];
		const uint2 q = ld_u2_sys((const uint2 *)seg_in.p[s] + j);
		const uint2 o = gh::search_one(table, g, q);
		((uint2 *)seg_out.p[s])[j] = o;                               /* local staging or the origin's, over NVLink */
	}
}

/* Tell every origin that its results are in place (fused path). */
__global__ void results_publish_kernel(int G, int my_rank, Ptrs peer_flag, uint32_t seq)
{
	int d = threadIdx.x;
	if (d >= G) return;
	__threadfence_system();
	asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)peer_flag.p[d] + my_rank), "r"(seq) : "memory");
}

/* ---- K3: results back into request order ---------------------------------------------------------- */
__global__ void __launch_bounds__(256)
route_gather_kernel(Ptrs staged /* [G] -> uint2[cap] */, const uint32_t *__restrict__ perm, const uint32_t *counts,
		size_t cap, int G, uint2 *__restrict__ out, const uint32_t *flags, uint32_t wait_seq, uint32_t *err)
{
	__shared__ uint32_t prefix[kMaxShards + 1];
	__shared__ int ok;
	if (threadIdx.x == 0) {
		ok = 1;
		if (wait_seq) for (int s = 0; s < G && ok; s++) ok = wait_flag(flags + s, wait_seq, 2000000000ULL, err);
		uint32_t acc = 0;
		for (int s = 0; s < G; s++) { prefix[s] = acc; acc += counts[s]; }
		prefix[G] = acc;
	}
	__syncthreads();
	if (!ok) return;
	const uint32_t total = prefix[G];
	int s = 0;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
		while (e >= prefix[s + 1]) s++;
		const uint32_t j = e - prefix[s];
		const uint2 v = ld_u2_sys((const uint2 *)staged.p[s] + j);
		gh::st_stream_u2(out + perm[(size_t)s * cap + j], v);
	}
}

__global__ void __launch_bounds__(256)
delete_segments_kernel(gh::Bucket *table, gh::Geom g, int G, Ptrs seg_in, const uint32_t *seg_count, gh::Stats *st)
{
	__shared__ uint32_t prefix[kMaxShards + 1];
	if (threadIdx.x == 0) {
		uint32_t acc = 0;
		for (int s = 0; s < G; s++) { prefix[s] = acc; acc += seg_count[s]; }
		prefix[G] = acc;
	}
	__syncthreads();
	const uint32_t total = prefix[G];
	int s = 0;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
		while (e >= prefix[s + 1]) s++;
		const uint32_t *p = (const uint32_t *)seg_in.p[s] + 3 * (size_t)(e - prefix[s]);
		int z = g.layout == gh::kLayoutPairs ? gh::delete_one<true>(table, g, p[0], p[1], p[2])
		                                     : gh::delete_one<false>(table, g, p[0], p[1], p[2]);
		if (st && z) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
	}
}

/* stream-ordered wait: everything enqueued after this kernel sees what the peers published before raising their flags */
__global__ void wait_flags_kernel(const uint32_t *flags, int G, uint32_t want, uint32_t *err)
{
	if (threadIdx.x < G) wait_flag(flags + threadIdx.x, want, 2000000000ULL, err);
}

/* ================= fused variants: 3 launches per routed search, 2 per routed insert/delete =================
 * Launch count, not bytes, is what a 64 K-request batch pays for (a graph kernel node costs ~1 us of front-end
 * time on B200), so publication and waiting are folded into the kernels that produce / consume the data:
 *   scatter+publish   every CTA first waits until the owners have consumed this rank's previous batch
 *                     (res_flag >= seq-1), scatters, and the LAST CTA to finish (ticket) publishes counts + flag
 *   serve             waits for all sources' flags, looks up / inserts / deletes what arrived, last CTA raises the
 *                     result flags on every origin
 *   gather            (searches) waits for the result flags, un-permutes                                        */

struct PubArgs {
	Ptrs peer_count, peer_flag;      /* [G] -> uint32[G] on each peer */
	uint32_t *ticket;                /* local, zero between launches */
	int my_rank;
	uint32_t seq;
};

__device__ __forceinline__ bool last_cta_done(uint32_t *ticket)
{
	__shared__ int last;
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence_system();                                          /* this CTA's peer stores before the ticket */
		last = atomicAdd(ticket, 1u) == gridDim.x - 1;
	}
	__syncthreads();
	return last != 0;
}

/* Slot allocation is warp-aggregated: the lanes of a warp that go to the same owner are found with
 * __match_any_sync, their leader takes one shared-memory atomic for all of them (G <= 8 atomics per warp instead
 * of 32 on one or two hot counters), one global atomic per (CTA, owner) reserves the CTA's run in the owner's
 * region, and the run is written in slot order so that the stores of neighbouring lanes coalesce -- over NVLink
 * when the owner is a peer.  kItems requests per thread and tile amortise the three barriers. */
template <int kWords, int kItems>
__global__ void __launch_bounds__(256)
route_scatter_pub_kernel(const uint32_t *__restrict__ in, size_t n, uint32_t hash_mask_total, int shift, int G,
		Ptrs dst, uint32_t *counts2 /* [2][8] */, uint32_t *perm, size_t cap, PubArgs pub,
		const uint32_t *ack_flags, uint32_t *err)
{
	__shared__ uint32_t blk_count[kMaxShards], blk_base[kMaxShards];
	__shared__ int ok;
	uint32_t *counts = counts2 + 8 * (pub.seq & 1u);
	const unsigned lane = threadIdx.x & 31u;
	if (threadIdx.x == 0) {
		ok = 1;
		if (ack_flags) for (int s = 0; s < G && ok; s++) ok = wait_flag(ack_flags + s, pub.seq - 1u, 2000000000ULL, err);
	}
	__syncthreads();
	if (ok) {
		const size_t tile_sz = (size_t)blockDim.x * kItems;
		for (size_t tile = (size_t)blockIdx.x * tile_sz; tile < n; tile += (size_t)gridDim.x * tile_sz) {
			if (threadIdx.x < kMaxShards) blk_count[threadIdx.x] = 0;
			__syncthreads();
			uint32_t w[kItems][kWords]; uint32_t d[kItems], rank[kItems];
#pragma unroll
			for (int it = 0; it < kItems; it++) {
				const size_t i = tile + (size_t)it * blockDim.x + threadIdx.x;
				const bool live = i < n;
				d[it] = 0xffu; rank[it] = 0;
				if (live) {
#pragma unroll
					for (int k = 0; k < kWords; k++) w[it][k] = gh::ld_stream_u32(in + kWords * i + k);
					d[it] = (w[it][1] & hash_mask_total) >> shift;
				}
				const unsigned peers = __match_any_sync(0xffffffffu, d[it]);       /* lanes with the same owner */
				const int leader = __ffs(peers) - 1;
				uint32_t base = 0;
				if (live && (int)lane == leader) base = atomicAdd(&blk_count[d[it]], (uint32_t)__popc(peers));
				base = __shfl_sync(0xffffffffu, base, leader);
				rank[it] = base + __popc(peers & ((1u << lane) - 1u));
			}
			__syncthreads();
			if (threadIdx.x < G && blk_count[threadIdx.x])
				blk_base[threadIdx.x] = atomicAdd(&counts[threadIdx.x], blk_count[threadIdx.x]);
			__syncthreads();
#pragma unroll
			for (int it = 0; it < kItems; it++) {
				const size_t i = tile + (size_t)it * blockDim.x + threadIdx.x;
				if (i < n) {
					const size_t slot = (size_t)blk_base[d[it]] + rank[it];
					uint32_t *q = (uint32_t *)dst.p[d[it]] + kWords * slot;
					if (kWords == 2) *reinterpret_cast<uint2 *>(q) = make_uint2(w[it][0], w[it][1]);
					else { q[0] = w[it][0]; q[1] = w[it][1]; q[kWords - 1] = w[it][kWords - 1]; }
					if (perm) perm[(size_t)d[it] * cap + slot] = (uint32_t)i;
				}
			}
			__syncthreads();
		}
	}
	if (last_cta_done(pub.ticket)) {
		if (threadIdx.x < G) {
			const int dd = threadIdx.x;
			const uint32_t c = ((volatile uint32_t *)counts)[dd];
			((volatile uint32_t *)pub.peer_count.p[dd])[pub.my_rank] = c;
			__threadfence_system();
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)pub.peer_flag.p[dd] + pub.my_rank), "r"(pub.seq) : "memory");
		}
		if (threadIdx.x < kMaxShards) counts2[8 * ((pub.seq + 1u) & 1u) + threadIdx.x] = 0;   /* next batch's counters */
		if (threadIdx.x == 0) *pub.ticket = 0;
	}
}

template <bool kPairs, int kOp /* 0 search, 1 insert, 2 delete */>
__global__ void __launch_bounds__(256)
serve_kernel(gh::Bucket *table, gh::Geom g, int G, Ptrs seg_in, const uint32_t *seg_count, Ptrs seg_out,
		const uint32_t *req_flags, uint32_t *err, PubArgs pub, gh::Stats *st)
{
	__shared__ uint32_t prefix[kMaxShards + 1];
	__shared__ int ok;
	if (threadIdx.x == 0) {
		ok = 1;
		if (req_flags) for (int s = 0; s < G && ok; s++) ok = wait_flag(req_flags + s, pub.seq, 2000000000ULL, err);
		uint32_t acc = 0;
		for (int s = 0; s < G; s++) { prefix[s] = acc; acc += ((const volatile uint32_t *)seg_count)[s]; }
		prefix[G] = acc;
	}
	__syncthreads();
	if (ok && kPairs && kOp != 0) {
		/* updates, pair layout: two lanes per request (gh::insert_pair / gh::delete_pair: the bucket is one L2 request);
		 * the trip count is warp-uniform, lanes past the end idle but take part in the shuffles */
		const uint32_t total = prefix[G], total_up = (total + 15u) & ~15u;
		const unsigned lane = threadIdx.x & 31u;
		int s = 0;
		for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 1; e < total_up; e += (gridDim.x * blockDim.x) >> 1) {
			const bool have = e < total;
			uint32_t a = 0, b = 0, c = 0;
			if (have) {
				while (e >= prefix[s + 1]) s++;
				/* the inbox was complete before this kernel was launched (flag wait in the stream): ordinary coalescing loads */
				const uint32_t *p = (const uint32_t *)seg_in.p[s] + 3 * (size_t)(e - prefix[s]);
				a = gh::ld_stream_u32(p); b = gh::ld_stream_u32(p + 1); c = gh::ld_stream_u32(p + 2);
			}
			if (kOp == 1) gh::insert_pair(table, g, have, a, b, c, st, lane);
			else {
				const int z = gh::delete_pair(table, g, have, a, b, c, lane);
				if (st && z && (lane & 1u) == 0) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
			}
		}
	} else if (ok) {
		const uint32_t total = prefix[G];
		int s = 0;
		for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
			while (e >= prefix[s + 1]) s++;
			const uint32_t j = e - prefix[s];
			if (kOp == 0) {
				const uint2 q = ld_u2_sys((const uint2 *)seg_in.p[s] + j);
				((uint2 *)seg_out.p[s])[j] = gh::search_one(table, g, q);
			} else {
				const uint32_t *p = (const uint32_t *)seg_in.p[s] + 3 * (size_t)j;
				uint2 a; uint32_t loc;                                   /* 12-byte records: scalar loads */
				asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(a.x) : "l"(p) : "memory");
				asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(a.y) : "l"(p + 1) : "memory");
				asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(loc) : "l"(p + 2) : "memory");
				if (kOp == 1) gh::insert_one<kPairs>(table, g, a.x, a.y, loc, st);
				else {
					int z = gh::delete_one<kPairs>(table, g, a.x, a.y, loc);
					if (st && z) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
				}
			}
		}
	}
	if (last_cta_done(pub.ticket)) {
		if (threadIdx.x < G)
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)pub.peer_flag.p[threadIdx.x] + pub.my_rank), "r"(pub.seq) : "memory");
		if (threadIdx.x == 0) *pub.ticket = 0;
	}
}

/* ================= tile-sorted routing: scatter and gather that move whole runs =================
 * route_scatter_pub_kernel stores each request from the lane that loaded it, so one warp store touches up to G
 * runs of ~4 requests (32 B pieces), and route_gather_kernel un-permutes with one random 8 B store per result:
 * about 0.5 and 1.1 L2 requests per search, on top of the 2.1 the lookup itself needs -- and requests are what the
 * memory system counts (profiles/r01_l2_requests.md).  Here a CTA sorts its tile of kRouteTile requests by owner in
 * shared memory first and writes each owner's run as one contiguous piece (128 requests = 1 KB on average at G = 8;
 * full 128 B lines locally, large NVLink packets to peers).  What the gather needs to undo it is tiny: per tile the
 * G run starts and lengths, per request its position in the sorted tile (16 bits).  The gather reads the G result
 * runs of a tile contiguously into shared memory and writes the tile's results in request order, coalesced.
 *   map layout: uint16 pos[cap_pad] | per tile { uint32 base[8], cnt[8] }          (cap_pad = cap rounded up to a tile) */
constexpr int kRouteTile = 1024;

__device__ __forceinline__ int run_of(const uint32_t *off, uint32_t q)      /* off[1..7]: run starts inside the sorted tile */
{
	return (int)(q >= off[1]) + (int)(q >= off[2]) + (int)(q >= off[3]) + (int)(q >= off[4])
	     + (int)(q >= off[5]) + (int)(q >= off[6]) + (int)(q >= off[7]);
}

template <int kWords>
__global__ void __launch_bounds__(256)
route_scatter_tiles_kernel(const uint32_t *__restrict__ in, size_t n, uint32_t hash_mask_total, int shift, int G,
		Ptrs dst, uint32_t *counts2 /* [2][8] */, uint16_t *pos, uint32_t *meta, PubArgs pub)
{
	constexpr int kItems = kRouteTile / 256;
	__shared__ uint32_t cnt[kMaxShards], base[kMaxShards], off[kMaxShards + 1];
	__shared__ uint32_t *dst_s[kMaxShards];
	__shared__ __align__(16) uint32_t stage[kRouteTile * kWords];
	uint32_t *counts = counts2 + 8 * (pub.seq & 1u);
	const unsigned lane = threadIdx.x & 31u;
	if (threadIdx.x < kMaxShards) dst_s[threadIdx.x] = (uint32_t *)dst.p[threadIdx.x];
	const size_t tiles = (n + kRouteTile - 1) / kRouteTile;
	for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		const size_t t0 = tile * kRouteTile;
		const uint32_t tile_n = (uint32_t)(n - t0 < (size_t)kRouteTile ? n - t0 : (size_t)kRouteTile);
		if (threadIdx.x < kMaxShards) cnt[threadIdx.x] = 0;
		__syncthreads();
		uint32_t w[kItems][kWords], d[kItems], rank[kItems];
#pragma unroll
		for (int it = 0; it < kItems; it++) {
			const uint32_t k = it * 256 + threadIdx.x;
			const bool live = k < tile_n;
			d[it] = 0xffu; rank[it] = 0;
			if (live) {
				if (kWords == 2) {
					const uint2 v = gh::ld_stream_u2((const uint2 *)in + t0 + k);
					w[it][0] = v.x; w[it][1] = v.y;
				} else {
#pragma unroll
					for (int j = 0; j < kWords; j++) w[it][j] = gh::ld_stream_u32(in + kWords * (t0 + k) + j);
				}
				d[it] = (w[it][1] & hash_mask_total) >> shift;           /* owner = top bits of bucket 1 (== of bucket 2) */
			}
			const unsigned peers = __match_any_sync(0xffffffffu, d[it]);
			const int leader = __ffs(peers) - 1;
			uint32_t b = 0;
			if (live && (int)lane == leader) b = atomicAdd(&cnt[d[it]], (uint32_t)__popc(peers));
			b = __shfl_sync(0xffffffffu, b, leader);
			rank[it] = b + __popc(peers & ((1u << lane) - 1u));
		}
		__syncthreads();
		if (threadIdx.x < kMaxShards) {
			const uint32_t c = cnt[threadIdx.x];
			base[threadIdx.x] = c ? atomicAdd(&counts[threadIdx.x], c) : 0u;   /* this tile's run in the owner's region */
			uint32_t o = 0;
			for (int j = 0; j < (int)threadIdx.x; j++) o += cnt[j];
			off[threadIdx.x] = o;
			if (threadIdx.x == kMaxShards - 1) off[kMaxShards] = o + c;
			if (meta) { meta[tile * 16 + threadIdx.x] = base[threadIdx.x]; meta[tile * 16 + 8 + threadIdx.x] = c; }
		}
		__syncthreads();
#pragma unroll
		for (int it = 0; it < kItems; it++) {
			const uint32_t k = it * 256 + threadIdx.x;
			if (k < tile_n) {
				const uint32_t p = off[d[it]] + rank[it];
#pragma unroll
				for (int j = 0; j < kWords; j++) stage[p * kWords + j] = w[it][j];
				if (pos) pos[t0 + k] = (uint16_t)p;
			}
		}
		__syncthreads();
		for (uint32_t q = threadIdx.x; q < tile_n; q += 256) {           /* sorted order: neighbours share a run */
			const int dd = run_of(off, q);
			uint32_t *o = dst_s[dd] + (size_t)kWords * (base[dd] + q - off[dd]);
			if (kWords == 2) *reinterpret_cast<uint2 *>(o) = make_uint2(stage[2 * q], stage[2 * q + 1]);
			else { o[0] = stage[3 * q]; o[1] = stage[3 * q + 1]; o[2] = stage[3 * q + 2]; }
		}
		__syncthreads();
	}
	if (last_cta_done(pub.ticket)) {
		if (threadIdx.x < G) {
			const int dd = threadIdx.x;
			const uint32_t c = ((volatile uint32_t *)counts)[dd];
			((volatile uint32_t *)pub.peer_count.p[dd])[pub.my_rank] = c;
			__threadfence_system();
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)pub.peer_flag.p[dd] + pub.my_rank), "r"(pub.seq) : "memory");
		}
		if (threadIdx.x < kMaxShards) counts2[8 * ((pub.seq + 1u) & 1u) + threadIdx.x] = 0;   /* next batch's counters */
		if (threadIdx.x == 0) *pub.ticket = 0;
	}
}

__global__ void __launch_bounds__(256)
route_gather_tiles_kernel(Ptrs staged /* [G] -> uint2[cap] */, const uint16_t *__restrict__ pos, const uint32_t *__restrict__ meta,
		uint2 *__restrict__ out, size_t n)
{
	__shared__ uint32_t base[kMaxShards], off[kMaxShards + 1];
	__shared__ const uint2 *src_s[kMaxShards];
	__shared__ __align__(16) uint2 stage[kRouteTile];
	if (threadIdx.x < kMaxShards) src_s[threadIdx.x] = (const uint2 *)staged.p[threadIdx.x];
	const size_t tiles = (n + kRouteTile - 1) / kRouteTile;
	for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		const size_t t0 = tile * kRouteTile;
		const uint32_t tile_n = (uint32_t)(n - t0 < (size_t)kRouteTile ? n - t0 : (size_t)kRouteTile);
		__syncthreads();                                                  /* previous tile's stage[] and off[] are done with */
		if (threadIdx.x < kMaxShards) {
			base[threadIdx.x] = meta[tile * 16 + threadIdx.x];
			uint32_t o = 0;
			for (int j = 0; j < (int)threadIdx.x; j++) o += meta[tile * 16 + 8 + j];
			off[threadIdx.x] = o;
			if (threadIdx.x == kMaxShards - 1) off[kMaxShards] = o + meta[tile * 16 + 8 + threadIdx.x];
		}
		__syncthreads();
		for (uint32_t q = threadIdx.x; q < tile_n; q += 256) {
			const int dd = run_of(off, q);
			stage[q] = ld_u2_sys(src_s[dd] + base[dd] + q - off[dd]);
		}
		__syncthreads();
		for (uint32_t k = threadIdx.x; k < tile_n; k += 256)
			gh::st_stream_u2(out + t0 + k, stage[pos[t0 + k]]);
	}
}

int fill_ptrs(Ptrs &P, const void *const *src, int G)
{
	if (G < 1 || G > kMaxShards || !src) return -1;
	for (int k = 0; k < kMaxShards; k++) P.p[k] = k < G ? (void *)src[k] : nullptr;
	return 0;
}

/* launch-shape knobs of the routed path, read once from the environment (experiments; defaults are what bench.py runs) */
int env_int(const char *name, int dflt)
{
	const char *e = getenv(name);
	if (!e || !*e) return dflt;
	int v = atoi(e);
	return v > 0 ? v : dflt;
}
int serve_staged(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("GPUHASH_SERVE_STAGED"); v = (e && e[0] == '0') ? 0 : 1; }
	return v;
}

unsigned grid_for(size_t n, int per_sm)
{
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	size_t blocks = (n + 255) / 256, cap = (size_t)sms * per_sm;
	if (blocks > cap) blocks = cap;
	return (unsigned)(blocks ? blocks : 1);
}

}  // namespace

extern "C" int gpuhash_route_scatter(const void *in_d, size_t n, int elem_words, uint32_t hash_mask_total,
		int log2_shards, const void *const *dst_ptrs /* HOST array [G] of device/peer pointers */,
		uint32_t *counts_d, uint32_t *perm_d, size_t cap, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs D;
	if (log2_shards < 0 || log2_shards > 3 || (elem_words != 2 && elem_words != 3) || fill_ptrs(D, dst_ptrs, G) || !counts_d || n > cap) return -1;
	cudaStream_t s = (cudaStream_t)stream;
	cudaError_t e = cudaMemsetAsync(counts_d, 0, sizeof(uint32_t) * kMaxShards, s);
	if (e != cudaSuccess) return (int)e;
	if (n == 0) return 0;
	int bits = 0; while ((hash_mask_total >> bits) & 1u) bits++;
	const int shift = bits - log2_shards;
	if (shift < 0) return -1;
	const unsigned blocks = grid_for(n, 16);
	if (elem_words == 2)
		route_scatter_kernel<2><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts_d, perm_d, cap);
	else
		route_scatter_kernel<3><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts_d, perm_d, cap);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_route_publish(const uint32_t *counts_d, int log2_shards, int my_rank,
		const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs, uint32_t seq, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs C, F;
	if (fill_ptrs(C, peer_count_ptrs, G) || fill_ptrs(F, peer_flag_ptrs, G)) return -1;
	route_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts_d, G, my_rank, C, F, seq);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_search_segments(const gpuhash_geom_t *g, const void *table_d, int num_seg,
		const void *const *seg_in_ptrs, const uint32_t *seg_count_d, const void *const *seg_out_ptrs,
		size_t max_total, const uint32_t *flags_d, uint32_t wait_seq, uint32_t *err_d, void *stream)
{
	Ptrs I, O;
	if (!g || fill_ptrs(I, seg_in_ptrs, num_seg) || fill_ptrs(O, seg_out_ptrs, num_seg) || !seg_count_d) return -1;
	if (wait_seq && (!flags_d || !err_d)) return -1;
	gh::Geom gg; gg.hash_mask = g->hash_mask; gg.block_mask = g->block_mask; gg.algo = g->algo; gg.max_cuckoo = g->max_cuckoo; gg.layout = g->layout;
	search_segments_kernel<<<grid_for(max_total, 8), 256, 0, (cudaStream_t)stream>>>((const gh::Bucket *)table_d, gg, num_seg, I,
			seg_count_d, O, flags_d, wait_seq, err_d);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_results_publish(int log2_shards, int my_rank, const void *const *peer_flag_ptrs, uint32_t seq, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs F;
	if (fill_ptrs(F, peer_flag_ptrs, G)) return -1;
	results_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(G, my_rank, F, seq);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_route_gather(const void *const *staged_ptrs, const uint32_t *perm_d, const uint32_t *counts_d,
		size_t cap, int log2_shards, void *out_d, size_t n, const uint32_t *flags_d, uint32_t wait_seq, uint32_t *err_d, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs S;
	if (fill_ptrs(S, staged_ptrs, G) || !perm_d || !counts_d || (n && !out_d)) return -1;
	if (wait_seq && (!flags_d || !err_d)) return -1;
	if (n == 0 && !wait_seq) return 0;                       /* with a flag wait the (empty) kernel still orders the stream */
	route_gather_kernel<<<grid_for(n, env_int("GPUHASH_GATHER_CTAS_PER_SM", 8)), 256, 0, (cudaStream_t)stream>>>(S, perm_d, counts_d, cap, G, (uint2 *)out_d, flags_d, wait_seq, err_d);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_delete_segments(const gpuhash_geom_t *g, void *table_d, int num_seg, const void *const *seg_in_ptrs,
		const uint32_t *seg_count_d, size_t max_total, gpuhash_stats_t *stats_d, void *stream)
{
	Ptrs I;
	if (!g || fill_ptrs(I, seg_in_ptrs, num_seg) || !seg_count_d) return -1;
	gh::Geom gg; gg.hash_mask = g->hash_mask; gg.block_mask = g->block_mask; gg.algo = g->algo; gg.max_cuckoo = g->max_cuckoo; gg.layout = g->layout;
	delete_segments_kernel<<<grid_for(max_total, 8), 256, 0, (cudaStream_t)stream>>>((gh::Bucket *)table_d, gg, num_seg, I,
			seg_count_d, (gh::Stats *)stats_d);
	return (int)cudaGetLastError();
}

/* Stream-ordered wait for flags_d[0..num) >= want.
 * Preferred: stream memory operations (cuStreamBatchMemOp, CU_STREAM_WAIT_VALUE_GEQ): the wait is done by the
 * front end, holds no SM and is captured into CUDA graphs as a mem-op node.  A waiting KERNEL, even one CTA,
 * blocks its hardware queue while it spins, and with many lanes in flight the lanes serialise behind each
 * other's waiters (2 GPUs, 32 lanes: 26 us per step however many lanes).  Fallback (GPUHASH_WAIT_MODE=kernel, or
 * a driver without the entry point): the one-CTA kernel with a 2 s timeout. */
typedef CUresult (*batch_memop_fn)(CUstream, unsigned int, CUstreamBatchMemOpParams *, unsigned int);
static batch_memop_fn g_batch_memop;
static int g_wait_mode = -1;          /* -1 unknown, 0 kernel, 1 mem-op */

static void resolve_wait_mode(void)
{
	const char *env = getenv("GPUHASH_WAIT_MODE");
	g_wait_mode = 0;
	if (env && !strcmp(env, "kernel")) return;
	void *fn = NULL;
	cudaDriverEntryPointQueryResult st;
	if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &fn, cudaEnableDefault, &st) == cudaSuccess && fn && st == cudaDriverEntryPointSuccess) {
		g_batch_memop = (batch_memop_fn)fn;
		g_wait_mode = 1;
	}
}

extern "C" int gpuhash_wait_flags(const uint32_t *flags_d, int num, uint32_t want, uint32_t *err_d, void *stream)
{
	if (!flags_d || !err_d || num < 1 || num > kMaxShards) return -1;
	if (g_wait_mode < 0) resolve_wait_mode();
	if (g_wait_mode == 1) {
		CUstreamBatchMemOpParams ops[kMaxShards];
		memset(ops, 0, sizeof ops);
		for (int k = 0; k < num; k++) {
			ops[k].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
			ops[k].waitValue.address = (CUdeviceptr)(uintptr_t)(flags_d + k);
			ops[k].waitValue.value = want;
			ops[k].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
		}
		CUresult r = g_batch_memop((CUstream)stream, (unsigned)num, ops, 0);
		if (r == CUDA_SUCCESS) return 0;
		g_wait_mode = 0;                                   /* not supported here (or not capturable): use the kernel */
	}
	wait_flags_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(flags_d, num, want, err_d);
	return (int)cudaGetLastError();
}
extern "C" int gpuhash_wait_mode(void) { if (g_wait_mode < 0) resolve_wait_mode(); return g_wait_mode; }

/* serve, op 0, four lanes per request: each bucket of a routed request is ONE L2 request (see search_quad_kernel).
 * A warp handles 8 consecutive requests of ONE source region (the request index space pads every region to a
 * multiple of 8), and their 8 results leave as two aligned 32 B stores by two lanes: the staging area is
 * peer-mapped memory, where ncu shows partial-sector stores costing a DRAM write each (73 B written per 8 B result
 * before this). */
template <bool kPairs>
__global__ void __launch_bounds__(256)
serve_search_quad_kernel(const gh::Bucket *__restrict__ table, gh::Geom g, int G, Ptrs seg_in, const uint32_t *seg_count, Ptrs seg_out,
		const uint32_t *req_flags, uint32_t *err, PubArgs pub)
{
	__shared__ uint32_t prefix[kMaxShards + 1], count[kMaxShards];
	__shared__ int ok;
	if (threadIdx.x == 0) {
		ok = 1;
		if (req_flags) for (int s = 0; s < G && ok; s++) ok = wait_flag(req_flags + s, pub.seq, 2000000000ULL, err);
		uint32_t acc = 0;
		for (int s = 0; s < G; s++) {
			const uint32_t c = ((const volatile uint32_t *)seg_count)[s];
			prefix[s] = acc; count[s] = c; acc += (c + 7u) & ~7u;
		}
		prefix[G] = acc;
	}
	__syncthreads();
	if (ok) {
		const unsigned lane = threadIdx.x & 31u, sub = lane & 3u, grp0 = lane & ~3u, half = sub & 1u;
		const uint32_t total = prefix[G];                                 /* multiple of 8 */
		const uint32_t per_iter = (gridDim.x * blockDim.x) >> 2;
		int s = 0;
		for (uint32_t e = (blockIdx.x * blockDim.x + threadIdx.x) >> 2; e < total; e += per_iter) {
			while (e >= prefix[s + 1]) s++;
			const uint32_t j = e - prefix[s];
			const bool live = j < count[s];
			uint2 q = make_uint2(0u, 0u);
			gh::Row r;
#pragma unroll
			for (int k = 0; k < 8; k++) r.w[k] = 0;
			if (live) {
				q = ld_u2_sys((const uint2 *)seg_in.p[s] + j);
				const uint32_t b = sub < 2 ? gh::bucket1(g, q.y) : gh::bucket2(g, q.y, q.x);
				r = gh::ld_row_ro(table[b].w + 8 * half);
			}
			uint2 o;
			if (kPairs) {
				uint32_t m = (r.w[0] == q.x ? 1u : 0u) | (r.w[2] == q.x ? 2u : 0u) | (r.w[4] == q.x ? 4u : 0u) | (r.w[6] == q.x ? 8u : 0u);
				const uint32_t loc = (m & 1u) ? r.w[1] : (m & 2u) ? r.w[3] : (m & 4u) ? r.w[5] : r.w[7];
				if (!live) m = 0;
				const unsigned hits = (__ballot_sync(0xffffffffu, m != 0) >> grp0) & 0xfu;
				const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 1u) ? 0 : 1));
				const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 4u) ? 2 : 3));
				o = make_uint2((hits & 3u) ? l0 : 0u, (hits & 12u) ? l1 : 0u);
			} else {
				const uint32_t m = live ? gh::eq_mask(r, q.x) : 0u;
				const uint32_t msig = __shfl_sync(0xffffffffu, m, grp0 + (sub & 2u));
				const int l = __ffs(msig | 0x100u) - 1 & 7;
				uint32_t loc = r.w[0];
				if (l == 1) loc = r.w[1];
				if (l == 2) loc = r.w[2];
				if (l == 3) loc = r.w[3];
				if (l == 4) loc = r.w[4];
				if (l == 5) loc = r.w[5];
				if (l == 6) loc = r.w[6];
				if (l == 7) loc = r.w[7];
				if (!msig) loc = 0;
				o = make_uint2(__shfl_sync(0xffffffffu, loc, grp0 + 1), __shfl_sync(0xffffffffu, loc, grp0 + 3));
			}
			/* request k of this warp (k = 0..7) has its result on lane 4k; lanes 0 and 1 collect four each */
			uint32_t v[8];
#pragma unroll
			for (int k = 0; k < 4; k++) {
				const int src = 16 * (int)(lane & 1u) + 4 * k;
				v[2 * k] = __shfl_sync(0xffffffffu, o.x, src);
				v[2 * k + 1] = __shfl_sync(0xffffffffu, o.y, src);
			}
			if (lane < 2) {
				const uint32_t j0 = (e - (lane >> 2)) - prefix[s] + 4 * lane;   /* e of lane 0/1 is the warp's first request */
				uint32_t *dstp = (uint32_t *)seg_out.p[s] + 2 * (size_t)j0;
				if (j0 + 4 <= count[s]) {
					asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(dstp),
						"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
				} else {
					for (int k = 0; k < 4; k++)
						if (j0 + k < count[s]) *reinterpret_cast<uint2 *>(dstp + 2 * k) = make_uint2(v[2 * k], v[2 * k + 1]);
				}
			}
		}
	}
	if (last_cta_done(pub.ticket)) {
		if (threadIdx.x < G)
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)pub.peer_flag.p[threadIdx.x] + pub.my_rank), "r"(pub.seq) : "memory");
		if (threadIdx.x == 0) *pub.ticket = 0;
	}
}

/* serve, op 0, staged: the four-lane probe of serve_search_quad_kernel with the inbox read and the result written in
 * 64-request tiles by bulk copies (see search_quad_staged_kernel): one elected thread pulls the next tile of an inbox
 * region into shared memory while this one is probed, and the 64 results leave as ONE 512 B bulk store into the origin's
 * staging area -- over NVLink when the origin is a peer, where 512 B transactions instead of 32 B ones matter.
 * Persistent CTAs (grid <= 8 per SM, tunable) walk the tiles of all G regions; a region's last, partial tile is done
 * with plain loads/stores.  Before the last CTA raises the result flags every CTA waits for its bulk stores to be
 * complete (not only read out) and orders them before the ticket. */
template <bool kPairs>
__global__ void __launch_bounds__(256)
serve_search_staged_kernel(const gh::Bucket *__restrict__ table, gh::Geom g, int G, Ptrs seg_in, const uint32_t *seg_count, Ptrs seg_out,
		const uint32_t *req_flags, uint32_t *err, PubArgs pub)
{
	__shared__ __align__(128) uint2 q_s[2][gh::kTileReq];
	__shared__ __align__(128) uint2 o_s[2][gh::kTileReq];
	__shared__ __align__(8) unsigned long long bar[2];
	__shared__ uint32_t tprefix[kMaxShards + 1], count[kMaxShards];
	__shared__ const uint2 *in_p[kMaxShards];
	__shared__ uint2 *out_p[kMaxShards];
	__shared__ int ok;
	if (threadIdx.x == 0) {
		ok = 1;
		if (req_flags) for (int s = 0; s < G && ok; s++) ok = wait_flag(req_flags + s, pub.seq, 2000000000ULL, err);
		uint32_t acc = 0;
		for (int s = 0; s < G; s++) {
			const uint32_t c = ((const volatile uint32_t *)seg_count)[s];
			tprefix[s] = acc; count[s] = c; acc += (c + gh::kTileReq - 1) / gh::kTileReq;
		}
		tprefix[G] = acc;
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gh::smem_u32(&bar[0])));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(gh::smem_u32(&bar[1])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (threadIdx.x < kMaxShards) {
		in_p[threadIdx.x] = (const uint2 *)seg_in.p[threadIdx.x];
		out_p[threadIdx.x] = (uint2 *)seg_out.p[threadIdx.x];
	}
	__syncthreads();
	bool store_pending = false;                                       /* thread 0 only */
	if (ok) {
		const unsigned lane = threadIdx.x & 31u, sub = lane & 3u, grp0 = lane & ~3u, half = sub & 1u;
		const unsigned quad = threadIdx.x >> 2;
		const uint32_t tiles = tprefix[G];
		/* (region, first request, full?) of tile t; `s` only moves forward */
		auto locate = [&](uint32_t t, int &s, uint32_t &j0, bool &full) {
			while (t >= tprefix[s + 1]) s++;
			j0 = (t - tprefix[s]) * gh::kTileReq;
			full = j0 + gh::kTileReq <= count[s];
		};
		auto issue_load = [&](int s, uint32_t j0, int stage) {
			const uint32_t b = gh::smem_u32(&bar[stage]);
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(gh::kTileReq * 8) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				:: "r"(gh::smem_u32(&q_s[stage][0])), "l"(in_p[s] + j0), "r"(gh::kTileReq * 8), "r"(b) : "memory");
		};
		uint32_t t = blockIdx.x, it = 0, phases = 0;
		int s = 0, sn = 0;
		uint32_t j0 = 0, jn = 0; bool full = false, fulln = false;
		if (t < tiles) {
			locate(t, s, j0, full);
			if (threadIdx.x == 0 && full) issue_load(s, j0, 0);
		}
		sn = s;
		for (; t < tiles; t += gridDim.x, it++) {
			const int stage = (int)(it & 1u);
			const uint32_t tn = t + gridDim.x;
			if (tn < tiles) {
				locate(tn, sn, jn, fulln);
				if (threadIdx.x == 0 && fulln) issue_load(sn, jn, stage ^ 1);
			}
			const uint32_t j = j0 + quad;
			const bool live = j < count[s];
			uint2 q = make_uint2(0u, 0u);
			if (full) {                                                  /* parity per stage: partial tiles do not use the barrier */
				gh::mbar_wait(gh::smem_u32(&bar[stage]), (phases >> stage) & 1u);
				phases ^= 1u << stage;
				q = q_s[stage][quad];
			} else if (live) {
				q = ld_u2_sys(in_p[s] + j);
			}
			uint32_t h1, h2;
			const uint2 o = gh::quad_probe<kPairs>(table, g, q, live, sub, grp0, half, h1, h2);
			if (live && sub == 0) {
				if (full) o_s[stage][quad] = o; else out_p[s][j] = o;
			}
			if (full) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
			if (threadIdx.x == 0 && store_pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
			__syncthreads();
			if (threadIdx.x == 0 && full) {
				asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
					:: "l"(out_p[s] + j0), "r"(gh::smem_u32(&o_s[stage][0])), "r"(gh::kTileReq * 8) : "memory");
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				store_pending = true;
			}
			s = sn; j0 = jn; full = fulln;
		}
	}
	if (threadIdx.x == 0 && store_pending) {
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       /* complete, not just read out: the flag follows */
		asm volatile("fence.proxy.async;" ::: "memory");
	}
	if (last_cta_done(pub.ticket)) {
		if (threadIdx.x < G)
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)pub.peer_flag.p[threadIdx.x] + pub.my_rank), "r"(pub.seq) : "memory");
		if (threadIdx.x == 0) *pub.ticket = 0;
	}
}

/* serve, op 0, every warp on its own (gh::warp_tile_search, the search kernel's shape): a warp reads a 64-request tile of an
 * inbox region as ONE 512 B access (the next tile already in flight), probes it with four table loads per lane in flight and
 * writes the 64 results as ONE 512 B access straight into the origin's staging area (over NVLink when the origin is a peer).
 * No shared-memory staging, no barrier, no bulk copy; 64 registers, so three CTAs of eight warps per SM leave room for the
 * scatter / gather kernels of the other lanes to run next to it (the staged kernel: one load per lane in flight, 15.4 Gops/s
 * alone on 2 GPUs; this shape is the one the 1-GPU search path measures at 20-21).  Tiles are dealt round-robin to the warps of
 * the (persistent) grid; the last CTA raises the result flags. */
template <bool kPairs>
__global__ void __launch_bounds__(256, 3)
serve_search_warp_kernel(const gh::Bucket *__restrict__ table, gh::Geom g, int G, Ptrs seg_in, const uint32_t *seg_count, Ptrs seg_out, PubArgs pub)
{
	__shared__ uint32_t tprefix[kMaxShards + 1], count[kMaxShards];
	__shared__ const uint2 *in_p[kMaxShards];
	__shared__ uint2 *out_p[kMaxShards];
	if (threadIdx.x == 0) {
		uint32_t acc = 0;
		for (int s = 0; s < kMaxShards; s++) {
			const uint32_t c = s < G ? ((const volatile uint32_t *)seg_count)[s] : 0u;
			tprefix[s] = acc; count[s] = c; acc += (c + gh::kTileReq - 1) / gh::kTileReq;
		}
		tprefix[kMaxShards] = acc;
	}
	if (threadIdx.x < kMaxShards) {
		in_p[threadIdx.x] = (const uint2 *)seg_in.p[threadIdx.x];
		out_p[threadIdx.x] = (uint2 *)seg_out.p[threadIdx.x];
	}
	__syncthreads();
	const unsigned lane = threadIdx.x & 31u;
	const uint32_t tiles = tprefix[kMaxShards];
	const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
	uint32_t h1 = 0, h2 = 0;
	auto locate = [&](uint32_t t, int &s, uint32_t &j0, uint32_t &valid) {           /* s only moves forward */
		while (s < kMaxShards - 1 && t >= tprefix[s + 1]) s++;
		j0 = (t - tprefix[s]) * gh::kTileReq;
		valid = min((uint32_t)gh::kTileReq, count[s] - j0);
	};
	int s = 0, sn = 0; uint32_t j0 = 0, valid = 0, jn = 0, validn = 0;
	uint32_t t = warp;
	uint4 v = make_uint4(0u, 0u, 0u, 0u);
	if (t < tiles) { locate(t, s, j0, valid); v = gh::warp_tile_load<false>(in_p[s] + j0, valid, lane); }
	sn = s;
	for (; t < tiles; t += warps) {
		const uint32_t tn = t + warps;
		uint4 vn = make_uint4(0u, 0u, 0u, 0u);
		if (tn < tiles) { locate(tn, sn, jn, validn); vn = gh::warp_tile_load<false>(in_p[sn] + jn, validn, lane); }
		gh::warp_tile_search<kPairs, false, false>(table, g, in_p[s] + j0, out_p[s] + j0, valid, true, v, lane, h1, h2);
		s = sn; j0 = jn; valid = validn; v = vn;
	}
	if (last_cta_done(pub.ticket)) {
		if (threadIdx.x < G)
			asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"((uint32_t *)pub.peer_flag.p[threadIdx.x] + pub.my_rank), "r"(pub.seq) : "memory");
		if (threadIdx.x == 0) *pub.ticket = 0;
	}
}

static int fill_pub(PubArgs &P, int G, int my_rank, const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs,
		uint32_t *ticket_d, uint32_t seq)
{
	if (!ticket_d || fill_ptrs(P.peer_flag, peer_flag_ptrs, G)) return -1;
	if (peer_count_ptrs) { if (fill_ptrs(P.peer_count, peer_count_ptrs, G)) return -1; }
	else for (int k = 0; k < kMaxShards; k++) P.peer_count.p[k] = nullptr;
	P.ticket = ticket_d; P.my_rank = my_rank; P.seq = seq;
	return 0;
}

extern "C" int gpuhash_route_scatter_pub(const void *in_d, size_t n, int elem_words, uint32_t hash_mask_total, int log2_shards,
		const void *const *dst_ptrs, uint32_t *counts2_d, uint32_t *perm_d, size_t cap, int my_rank,
		const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs, uint32_t *ticket_d, uint32_t seq,
		const uint32_t *ack_flags_d, uint32_t *err_d, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs D; PubArgs P;
	if (log2_shards < 0 || log2_shards > 3 || (elem_words != 2 && elem_words != 3) || fill_ptrs(D, dst_ptrs, G) || !counts2_d
			|| n > cap || (ack_flags_d && !err_d) || !peer_count_ptrs || fill_pub(P, G, my_rank, peer_count_ptrs, peer_flag_ptrs, ticket_d, seq)) return -1;
	int bits = 0; while ((hash_mask_total >> bits) & 1u) bits++;
	const int shift = bits - log2_shards;
	if (shift < 0) return -1;
	cudaStream_t s = (cudaStream_t)stream;
	if (n <= ((size_t)1 << 17)) {                     /* small batch: one request per thread, as many CTAs as possible */
		const unsigned blocks = grid_for(n ? n : 1, env_int("GPUHASH_SCATTER_CTAS_PER_SM", 16));
		if (elem_words == 2)
			route_scatter_pub_kernel<2, 1><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts2_d, perm_d, cap, P, ack_flags_d, err_d);
		else
			route_scatter_pub_kernel<3, 1><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts2_d, perm_d, cap, P, ack_flags_d, err_d);
	} else {
		const unsigned blocks = grid_for((n + 3) / 4, env_int("GPUHASH_SCATTER_CTAS_PER_SM", 16));
		if (elem_words == 2)
			route_scatter_pub_kernel<2, 4><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts2_d, perm_d, cap, P, ack_flags_d, err_d);
		else
			route_scatter_pub_kernel<3, 4><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts2_d, perm_d, cap, P, ack_flags_d, err_d);
	}
	return (int)cudaGetLastError();
}

/* tile-sorted scatter + publish (map_d = NULL for insert/delete batches) and its gather; see kRouteTile above */
extern "C" size_t gpuhash_route_map_bytes(size_t cap)
{
	const size_t tiles = (cap + kRouteTile - 1) / kRouteTile;
	return tiles * kRouteTile * sizeof(uint16_t) + tiles * 16 * sizeof(uint32_t);
}

extern "C" int gpuhash_route_scatter_tiles(const void *in_d, size_t n, int elem_words, uint32_t hash_mask_total, int log2_shards,
		const void *const *dst_ptrs, uint32_t *counts2_d, void *map_d, size_t cap, int my_rank,
		const void *const *peer_count_ptrs, const void *const *peer_flag_ptrs, uint32_t *ticket_d, uint32_t seq, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs D; PubArgs P;
	if (log2_shards < 0 || log2_shards > 3 || (elem_words != 2 && elem_words != 3) || fill_ptrs(D, dst_ptrs, G) || !counts2_d
			|| n > cap || (n && !in_d) || !peer_count_ptrs || fill_pub(P, G, my_rank, peer_count_ptrs, peer_flag_ptrs, ticket_d, seq)) return -1;
	int bits = 0; while ((hash_mask_total >> bits) & 1u) bits++;
	const int shift = bits - log2_shards;
	if (shift < 0) return -1;
	const size_t cap_tiles = (cap + kRouteTile - 1) / kRouteTile;
	uint16_t *pos = (uint16_t *)map_d;
	uint32_t *meta = map_d ? (uint32_t *)((char *)map_d + cap_tiles * kRouteTile * sizeof(uint16_t)) : nullptr;
	size_t tiles = (n + kRouteTile - 1) / kRouteTile;
	const unsigned blocks = grid_for((tiles ? tiles : 1) * 256, env_int("GPUHASH_SCATTER_CTAS_PER_SM", 2));
	cudaStream_t s = (cudaStream_t)stream;
	if (elem_words == 2)
		route_scatter_tiles_kernel<2><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts2_d, pos, meta, P);
	else
		route_scatter_tiles_kernel<3><<<blocks, 256, 0, s>>>((const uint32_t *)in_d, n, hash_mask_total, shift, G, D, counts2_d, pos, meta, P);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_route_gather_tiles(const void *const *staged_ptrs, const void *map_d, size_t cap, int log2_shards,
		void *out_d, size_t n, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs S;
	if (log2_shards < 0 || log2_shards > 3 || fill_ptrs(S, staged_ptrs, G) || !map_d || n > cap || (n && !out_d)) return -1;
	if (n == 0) return 0;
	const size_t cap_tiles = (cap + kRouteTile - 1) / kRouteTile;
	const uint16_t *pos = (const uint16_t *)map_d;
	const uint32_t *meta = (const uint32_t *)((const char *)map_d + cap_tiles * kRouteTile * sizeof(uint16_t));
	const size_t tiles = (n + kRouteTile - 1) / kRouteTile;
	route_gather_tiles_kernel<<<grid_for(tiles * 256, env_int("GPUHASH_GATHER_CTAS_PER_SM", 2)), 256, 0, (cudaStream_t)stream>>>(S, pos, meta, (uint2 *)out_d, n);
	return (int)cudaGetLastError();
}

/* op: 0 search (seg_out_ptrs = origins' staging regions), 1 insert, 2 delete (seg_out_ptrs = NULL) */
extern "C" int gpuhash_serve(const gpuhash_geom_t *g, void *table_d, int op, int log2_shards, const void *const *seg_in_ptrs,
		const uint32_t *seg_count_d, const void *const *seg_out_ptrs, size_t max_total, const uint32_t *req_flags_d, uint32_t *err_d,
		int my_rank, const void *const *peer_res_flag_ptrs, uint32_t *ticket_d, uint32_t seq, gpuhash_stats_t *stats_d, void *stream)
{
	const int G = 1 << log2_shards;
	Ptrs I, O; PubArgs P;
	if (!g || op < 0 || op > 2 || fill_ptrs(I, seg_in_ptrs, G) || !seg_count_d || (req_flags_d && !err_d)
			|| fill_pub(P, G, my_rank, NULL, peer_res_flag_ptrs, ticket_d, seq)) return -1;
	if (op == 0) { if (fill_ptrs(O, seg_out_ptrs, G)) return -1; }
	else for (int k = 0; k < kMaxShards; k++) O.p[k] = nullptr;
	gh::Geom gg; gg.hash_mask = g->hash_mask; gg.block_mask = g->block_mask; gg.algo = g->algo; gg.max_cuckoo = g->max_cuckoo; gg.layout = g->layout;
	const unsigned blocks = grid_for((max_total ? max_total : 1) * 2, 8);      /* updates: two lanes per request (pair layout) */
	cudaStream_t s = (cudaStream_t)stream;
	gh::Bucket *t = (gh::Bucket *)table_d; gh::Stats *st = (gh::Stats *)stats_d;
	const bool pairs = gg.layout == gh::kLayoutPairs;
#define GH_SERVE(P_, OP_) serve_kernel<P_, OP_><<<blocks, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, req_flags_d, err_d, P, st)
	if (op == 0) {
		static int warp_mode = -1;                        /* GPUHASH_SERVE_MODE=warp | staged (default, until measured otherwise) */
		if (warp_mode < 0) { const char *e = getenv("GPUHASH_SERVE_MODE"); warp_mode = e && !strcmp(e, "warp") ? 1 : 0; }
		if (warp_mode) {
			const unsigned qb = grid_for(((max_total ? max_total : 1) + 63) / 64 * 32 + (size_t)G * 32, env_int("GPUHASH_SERVE_CTAS_PER_SM", 3));
			if (pairs) serve_search_warp_kernel<true><<<qb, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, P);
			else       serve_search_warp_kernel<false><<<qb, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, P);
		} else if (serve_staged()) {
			/* one tile of 64 requests per CTA and iteration; G partial tiles at most on top of max_total / 64 */
			const unsigned qb = grid_for(((max_total ? max_total : 1) + 63) / 64 * 256 + (size_t)G * 256, env_int("GPUHASH_SERVE_CTAS_PER_SM", 4));
			if (pairs) serve_search_staged_kernel<true><<<qb, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, req_flags_d, err_d, P);
			else       serve_search_staged_kernel<false><<<qb, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, req_flags_d, err_d, P);
		} else {
			const unsigned qb = grid_for((max_total ? max_total : 1) * 4, env_int("GPUHASH_SERVE_CTAS_PER_SM", 16));
			if (pairs) serve_search_quad_kernel<true><<<qb, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, req_flags_d, err_d, P);
			else       serve_search_quad_kernel<false><<<qb, 256, 0, s>>>(t, gg, G, I, seg_count_d, O, req_flags_d, err_d, P);
		}
	}
	else if (op == 1) { if (pairs) GH_SERVE(true, 1); else GH_SERVE(false, 1); }
	else              { if (pairs) GH_SERVE(true, 2); else GH_SERVE(false, 2); }
#undef GH_SERVE
	return (int)cudaGetLastError();
}

/* ---- CUDA IPC plumbing for the fused path (one process per GPU) ---- */
extern "C" int gpuhash_ipc_export(void *dev_ptr, void *handle_out_64B)
{
	return (int)cudaIpcGetMemHandle((cudaIpcMemHandle_t *)handle_out_64B, dev_ptr);
}
extern "C" void *gpuhash_ipc_import(const void *handle_64B)
{
	void *p = NULL;
	cudaIpcMemHandle_t h;
	memcpy(&h, handle_64B, sizeof h);
	return cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? p : NULL;
}
extern "C" int gpuhash_ipc_close(void *imported_ptr) { return (int)cudaIpcCloseMemHandle(imported_ptr); }
