/*
 * gpuhash_kernels.cuh -- device side of the B200-native Mega-KV hash index.
 *
 * Written from scratch for sm_100a; it is not a port of the reference kernels
 * (pzrq/megakv libgpuhash/gpu_hash.cu).  What it keeps is the *semantics* of
 * that file, cited per function, and the table bytes (bucket_t, gpu_hash.h:
 * 79-82): a 64 B bucket = one 32 B signature sector + one 32 B location sector.
 *
 * Work decomposition (DESIGN.md "Kernels"):
 *   reference: 8 lanes per request, one 4 B load per lane, ballot, __syncthreads
 *   here:      the whole 32 B signature row of a bucket is ONE 256-bit load
 *              (LDG.E.256, new on sm_100), so the "cooperative group per
 *              bucket" collapses into one thread that holds the row in eight
 *              registers and matches it with eight compares.  A warp therefore
 *              has 32 requests x 2 buckets = 64 independent sector reads in
 *              flight per load pair instead of 4, which is what a random-access
 *              HBM-bound kernel needs (Little's law: ~1e5 sectors in flight).
 *   conflicts: the reference arbitrates slot claims with "store, __syncthreads,
 *              re-read" inside one CUDA block and not at all across blocks;
 *              here every claim/eviction is an atomicCAS on the signature word,
 *              which is the linearisation point of the request.
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace gh {

constexpr int kSlots = 8;                 // ELEM_NUM          gpu_hash.h:49
constexpr uint32_t kAlgoCuckoo  = 0;      // HASH_CUCKOO       gpu_hash.h:73
constexpr uint32_t kAlgo2Choice = 1;      // HASH_2CHOICE      gpu_hash.h:72

struct Geom {                             // == gpuhash_geom_t (gpuhash_ex.h)
	uint32_t hash_mask;                   // buckets of THIS table - 1   (HASH_MASK, gpu_hash.h:61)
	uint32_t block_mask;                  // BLOCK_HASH_MASK of the logical table (gpu_hash.h:69)
	uint32_t algo;
	uint32_t max_cuckoo;                  // MAX_CUCKOO_NUM    gpu_hash.h:75
};

struct __align__(64) Bucket {             // bucket_t, gpu_hash.h:79-82
	uint32_t sig[kSlots];
	uint32_t loc[kSlots];
};

struct __align__(32) Row { uint32_t w[kSlots]; };

struct Stats {                            // == gpuhash_stats_t (gpuhash_ex.h)
	unsigned long long ins_skipped, ins_updated, ins_placed_b1, ins_placed_b2, ins_to_b2,
	                   ins_displaced, ins_dropped, ins_overwritten, ins_cas_retry, ins_gave_up,
	                   chain_hist[8],
	                   del_zeroed, del_requests_hit,
	                   search_hits_b1, search_hits_b2;
};

/* ------------------------------------------------------------------ memory ops */

// Table rows during search.  Random rows have no reuse inside a launch, so they are not
// allocated in L1.  Deliberately NOT .nc: kernels of other streams may be inserting into the
// same table (mega_scheduler.c runs one stream per worker with no cross-stream order), and a
// weak load that returns the old or the new word is what the reference's plain loads do too.
__device__ __forceinline__ Row ld_row_ro(const uint32_t* p)
{
	Row r;
	asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
		  "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
	return r;
}

// Same, plus an L2 prefetch hint of the enclosing 64 B: pulls the bucket's location sector
// into L2 together with its signature sector (used by the speculative search variant).
__device__ __forceinline__ Row ld_row_ro_pf64(const uint32_t* p)
{
	Row r;
	asm volatile("ld.global.L1::no_allocate.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
		  "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
	return r;
}

// Coherent at L2 (never served from L1): rows read by insert/delete, which race with
// other threads' CAS on the same words.
__device__ __forceinline__ Row ld_row_strong(const uint32_t* p)
{
	Row r;
	asm volatile("ld.relaxed.gpu.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
		  "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p) : "memory");
	return r;
}

__device__ __forceinline__ uint32_t ld_u32_ro(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

__device__ __forceinline__ void st_u32_strong(uint32_t* p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// request stream in / result stream out: touched once -> .cs (streaming, evict-first) so they
// do not displace table sectors from L2 when the table is L2-resident
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p)
{
	uint2 v;
	asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];"
		: "=r"(v.x), "=r"(v.y) : "l"(p));
	return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.global.cs.u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_stream_u2(uint2* p, uint2 v)
{
	asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};"
		:: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

/* ------------------------------------------------------------------ geometry */

// gpu_hash.cu:55
__device__ __forceinline__ uint32_t bucket1(const Geom& g, uint32_t hash)
{
	return hash & g.hash_mask;
}
// gpu_hash.cu:66-67, 171-172, 334-335, 471-472
__device__ __forceinline__ uint32_t bucket2(const Geom& g, uint32_t hash, uint32_t sig)
{
	return (((hash ^ sig) & g.block_mask) | (hash & ~g.block_mask)) & g.hash_mask;
}

// bit l set <=> row.w[l] == v   (the ballot of the reference, held by one thread)
__device__ __forceinline__ uint32_t eq_mask(const Row& r, uint32_t v)
{
	uint32_t m = 0;
#pragma unroll
	for (int l = 0; l < kSlots; l++) m |= (r.w[l] == v ? 1u : 0u) << l;
	return m;
}

// first set bit of `mask` (8 bits) scanning m, m+1, .., 7, 0, .., m-1: the "major location"
// rotation of gpu_hash.cu:139-147 / 301-309.  mask != 0.
__device__ __forceinline__ int first_from(uint32_t mask, int m)
{
	uint32_t rot = ((mask >> m) | (mask << (kSlots - m))) & 0xffu;
	return (__ffs(rot) - 1 + m) & (kSlots - 1);
}

/* ------------------------------------------------------------------ search */

// gpu_hash.cu:47-72 for one request.  Both buckets are always probed (the early exit at
// :61-63 is commented out in the reference).  With several matching slots the LOWEST one is
// reported: the reference lets all matching lanes store to the same word, and run on a B200 its
// kernel keeps the lowest lane's store (tests/golden/ref_search_cuckoo_16.npz, dup_* arrays).
template <bool kPrefetchLoc>
__device__ __forceinline__ void search_issue(const Bucket* __restrict__ table, const Geom& g,
		uint2 q /* x = sig, y = hash */, uint32_t& b1, uint32_t& b2, Row& r1, Row& r2)
{
	b1 = bucket1(g, q.y);
	b2 = bucket2(g, q.y, q.x);
	if (kPrefetchLoc) { r1 = ld_row_ro_pf64(table[b1].sig); r2 = ld_row_ro_pf64(table[b2].sig); }
	else              { r1 = ld_row_ro(table[b1].sig);      r2 = ld_row_ro(table[b2].sig); }
}

__device__ __forceinline__ uint2 search_finish(const Bucket* __restrict__ table, uint2 q,
		uint32_t b1, uint32_t b2, const Row& r1, const Row& r2)
{
	uint32_t m1 = eq_mask(r1, q.x), m2 = eq_mask(r2, q.x);
	uint2 o = make_uint2(0u, 0u);
	// the two location loads are independent: issue both before either is consumed
	const uint32_t* p1 = &table[b1].loc[__ffs(m1 | 0x100u) - 1 & 7];
	const uint32_t* p2 = &table[b2].loc[__ffs(m2 | 0x100u) - 1 & 7];
	uint32_t v1 = 0, v2 = 0;
	if (m1) v1 = ld_u32_ro(p1);
	if (m2) v2 = ld_u32_ro(p2);
	o.x = v1; o.y = v2;
	return o;
}

#ifdef GH_DEFINE_KERNELS   /* __global__ definitions: only libgpuhash.cu instantiates them */
// One thread per request, kQpt requests per thread issued back to back so that each thread
// keeps 2*kQpt sector reads in flight.  `out` gets both words of every request (0 = miss):
// the caller's cudaMemset of `out` (mega_scheduler.c:406) is fused away.
template <int kQpt, bool kPrefetchLoc>
__global__ void __launch_bounds__(256)
search_kernel(const uint2* __restrict__ in, uint2* __restrict__ out,
		const Bucket* __restrict__ table, size_t n, Geom g, Stats* st)
{
	const size_t tile = (size_t)blockDim.x * kQpt;
	for (size_t base = (size_t)blockIdx.x * tile; base < n; base += (size_t)gridDim.x * tile) {
		uint2 q[kQpt]; uint32_t b1[kQpt], b2[kQpt]; Row r1[kQpt], r2[kQpt];
		bool live[kQpt];
#pragma unroll
		for (int k = 0; k < kQpt; k++) {
			size_t i = base + (size_t)k * blockDim.x + threadIdx.x;
			live[k] = i < n;
			if (live[k]) q[k] = ld_stream_u2(in + i);
		}
#pragma unroll
		for (int k = 0; k < kQpt; k++)
			if (live[k]) search_issue<kPrefetchLoc>(table, g, q[k], b1[k], b2[k], r1[k], r2[k]);
#pragma unroll
		for (int k = 0; k < kQpt; k++) {
			if (!live[k]) continue;
			size_t i = base + (size_t)k * blockDim.x + threadIdx.x;
			uint2 o = search_finish(table, q[k], b1[k], b2[k], r1[k], r2[k]);
			st_stream_u2(out + i, o);
			if (st) {
				if (eq_mask(r1[k], q[k].x)) atomicAdd(&st->search_hits_b1, 1ULL);
				if (eq_mask(r2[k], q[k].x)) atomicAdd(&st->search_hits_b2, 1ULL);
			}
		}
	}
}

#endif  /* GH_DEFINE_KERNELS */

/* ------------------------------------------------------------------ delete */

// gpu_hash.cu:454-477 for one request: zero the signature of every slot whose signature AND
// location match; visit bucket 2 only if this request zeroed nothing in bucket 1.  The
// zeroing is a CAS(sig -> 0), so two identical requests of one batch behave like the
// sequential run: the first zeroes, the second sees a miss and goes on to bucket 2.
__device__ __forceinline__ int delete_in_bucket(Bucket* bk, uint32_t sig, uint32_t loc)
{
	Row s = ld_row_strong(bk->sig);
	uint32_t m = eq_mask(s, sig);
	if (!m) return 0;
	Row l = ld_row_strong(bk->loc);
	m &= eq_mask(l, loc);
	int zeroed = 0;
	while (m) {
		int slot = __ffs(m) - 1; m &= m - 1;
		if (atomicCAS(&bk->sig[slot], sig, 0u) == sig) zeroed++;
	}
	return zeroed;
}

__device__ __forceinline__ int delete_one(Bucket* table, const Geom& g,
		uint32_t sig, uint32_t hash, uint32_t loc)
{
	int z = delete_in_bucket(table + bucket1(g, hash), sig, loc);
	if (z) return z;                                                   // :465-468
	return delete_in_bucket(table + bucket2(g, hash, sig), sig, loc);
}

#ifdef GH_DEFINE_KERNELS
__global__ void __launch_bounds__(256)
delete_kernel(const uint32_t* __restrict__ in /* delem_t[n] as words */, Bucket* table,
		size_t n, Geom g, Stats* st)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
			i += (size_t)gridDim.x * blockDim.x) {
		uint32_t sig = ld_stream_u32(in + 3 * i), hash = ld_stream_u32(in + 3 * i + 1),
		         loc = ld_stream_u32(in + 3 * i + 2);
		int z = delete_one(table, g, sig, hash, loc);
		if (st && z) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
	}
}

#endif  /* GH_DEFINE_KERNELS */

/* ------------------------------------------------------------------ insert */

#define GH_COUNT(field) do { if (st) atomicAdd(&st->field, 1ULL); } while (0)

// gpu_hash.cu:256-430 (cuckoo) and :97-226 (2-choice) for one request, as a bounded
// lock-free loop.  One iteration = "read the signature row of the current bucket, decide,
// commit with one CAS":
//   signature present  -> store loc (update in place)             :277-287, 339-349
//   empty slot         -> CAS(sig[l]: 0 -> sig), then store loc   :303-327, 351-395
//                         l = first empty from the major location (sig0 & 7)
//   bucket 1 full      -> go to the alternate bucket              :330-336
//   alternate full     -> cuckoo:  victim slot sig0 & 7; CAS(sig[l]: victim -> sig),
//                                  exchange loc, carry the victim on -- with the REQUEST's
//                                  hash, never the victim's (:334-335, 403-404) -- at most
//                                  max_cuckoo times, then overwrite without re-homing (:414-422)
//                         2-choice: store sig into slot sig & 7, loc untouched (:197-209)
// A failed CAS means another request changed that slot first: the row is read again and the
// decision retaken, which is the order "other request, then this one" of a sequential run.
// Every failed CAS is another request's success, so the system as a whole always advances
// (lock-free); kMaxSteps additionally bounds one request's own loop.  Reaching it would need
// that many other requests to beat this one to the same slots; it is counted in ins_gave_up
// and the tests assert 0 even under 40 000 requests aimed at 64 buckets.
constexpr int kMaxSteps = 1 << 16;

__device__ __forceinline__ void insert_one(Bucket* table, const Geom& g,
		uint32_t sig0, uint32_t hash, uint32_t loc0, Stats* st)
{
	if (sig0 == 0 && loc0 == 0) { GH_COUNT(ins_skipped); return; }   // :101-104, 259-262

	uint32_t sig = sig0, loc = loc0;
	const int major = (int)(sig0 & (kSlots - 1));                    // ml_mask :139, 301
	uint32_t b = bucket1(g, hash);
	bool alt = false;
	uint32_t c = 0;                                                  // cuckoo_num :331

	for (int step = 0; step < kMaxSteps; step++) {
		Bucket* bk = table + b;
		Row r = ld_row_strong(bk->sig);
		uint32_t hit = eq_mask(r, sig);
		if (hit) {                                                   // update in place
			st_u32_strong(&bk->loc[__ffs(hit) - 1], loc);
			GH_COUNT(ins_updated);
			goto done;
		}
		uint32_t empty = eq_mask(r, 0u);
		if (empty) {
			int l = first_from(empty, major);
			uint32_t old = atomicCAS(&bk->sig[l], 0u, sig);
			if (old == 0u || old == sig) {                           // claimed, or a twin claimed it
				st_u32_strong(&bk->loc[l], loc);
				if (old == 0u) { if (alt) GH_COUNT(ins_placed_b2); else GH_COUNT(ins_placed_b1); }
				else GH_COUNT(ins_updated);
				goto done;
			}
			GH_COUNT(ins_cas_retry);
			continue;                                                // slot taken: look again
		}
		if (!alt) {                                                  // bucket 1 full
			alt = true;
			GH_COUNT(ins_to_b2);
			b = bucket2(g, hash, sig);
			continue;
		}
		// alternate bucket full
		const int l = (int)(sig0 & (kSlots - 1));                    // elem->sig :200, 360
		if (g.algo == kAlgo2Choice) {
			st_u32_strong(&bk->sig[l], sig);                         // loc NOT written :197-209
			GH_COUNT(ins_overwritten);
			goto done;
		}
		uint32_t vsig = r.w[l];
		if (atomicCAS(&bk->sig[l], vsig, sig) != vsig) { GH_COUNT(ins_cas_retry); continue; }
		uint32_t vloc = atomicExch(&bk->loc[l], loc);
		if (c < g.max_cuckoo) {                                      // :361-365, 397-405
			c++;
			GH_COUNT(ins_displaced);
			sig = vsig; loc = vloc;
			b = bucket2(g, hash, sig);                               // request's hash, victim's sig
			continue;
		}
		GH_COUNT(ins_dropped);                                       // :414-422
		goto done;
	}
	GH_COUNT(ins_gave_up);
done:
	if (st && g.algo == kAlgoCuckoo) atomicAdd(&st->chain_hist[c < 7 ? c : 7], 1ULL);
}

#ifdef GH_DEFINE_KERNELS
// The legacy entry point only knows num_blks on the host; segment sizes live in device
// memory (mega_scheduler.c:493-494).  So the grid is count-independent: every CTA builds
// the prefix sum of the segment sizes in shared memory and the grid strides over the
// concatenation.  Segment membership carries no meaning.
constexpr int kMaxSegChunk = 1024;

__global__ void __launch_bounds__(256)
insert_segments_kernel(Bucket* table, const uint32_t* const* __restrict__ blk_input,
		const int* __restrict__ blk_elem_num, int num_blks, Geom g, Stats* st)
{
	__shared__ unsigned long long prefix[kMaxSegChunk + 1];
	__shared__ const uint32_t* base[kMaxSegChunk];
	for (int seg0 = 0; seg0 < num_blks; seg0 += kMaxSegChunk) {
		int nseg = min(kMaxSegChunk, num_blks - seg0);
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned long long acc = 0;
			for (int k = 0; k < nseg; k++) {
				prefix[k] = acc;
				int c = blk_elem_num[seg0 + k];
				acc += c > 0 ? (unsigned long long)c : 0ULL;
			}
			prefix[nseg] = acc;
		}
		for (int k = threadIdx.x; k < nseg; k += blockDim.x) base[k] = blk_input[seg0 + k];
		__syncthreads();
		const unsigned long long total = prefix[nseg];
		int k = 0;
		for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
				e < total; e += (unsigned long long)gridDim.x * blockDim.x) {
			while (e >= prefix[k + 1]) k++;                          // e only grows
			const uint32_t* p = base[k] + 3 * (e - prefix[k]);
			insert_one(table, g, ld_stream_u32(p), ld_stream_u32(p + 1), ld_stream_u32(p + 2), st);
		}
	}
}

// host-known count (extended API, pipeline)
__global__ void __launch_bounds__(256)
insert_flat_kernel(Bucket* table, const uint32_t* __restrict__ in, size_t n, Geom g, Stats* st)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
			i += (size_t)gridDim.x * blockDim.x)
		insert_one(table, g, ld_stream_u32(in + 3 * i), ld_stream_u32(in + 3 * i + 1),
				ld_stream_u32(in + 3 * i + 2), st);
}

// Sequential execution of the SAME device code by one thread, segments in order, requests in
// order: reproduces the oracle slot for slot at any load factor.  Used by the parity tests
// (GPUHASH_INSERT_SERIAL); not a fast path.
__global__ void insert_serial_kernel(Bucket* table, const uint32_t* const* blk_input,
		const int* blk_elem_num, int num_blks, const uint32_t* flat, size_t flat_n, Geom g, Stats* st)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (flat) {
		for (size_t i = 0; i < flat_n; i++)
			insert_one(table, g, flat[3 * i], flat[3 * i + 1], flat[3 * i + 2], st);
		return;
	}
	for (int k = 0; k < num_blks; k++) {
		const uint32_t* p = blk_input[k];
		for (int i = 0; i < blk_elem_num[k]; i++)
			insert_one(table, g, p[3 * i], p[3 * i + 1], p[3 * i + 2], st);
	}
}

__global__ void delete_serial_kernel(const uint32_t* in, Bucket* table, size_t n, Geom g, Stats* st)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	for (size_t i = 0; i < n; i++) {
		int z = delete_one(table, g, in[3 * i], in[3 * i + 1], in[3 * i + 2]);
		if (st && z) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
	}
}

#endif  /* GH_DEFINE_KERNELS */

}  // namespace gh

#ifdef GH_DEFINE_KERNELS
/* ------------------------------------------------------------------ alternative search shape */

namespace gh {

__device__ __forceinline__ uint4 ld_half_row(const uint32_t* p)
{
	uint4 v;
	asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// Cooperative variant kept for comparison (tools/sweep.py, DESIGN.md "why one thread per request"):
// four lanes per request -- lanes 0,1 take the two 16 B halves of bucket 1's signature row, lanes
// 2,3 those of bucket 2 -- 128-bit loads, ballot inside the 4-lane group, the hit lane fetches the
// location, lane 0 stores the pair.  Same results as search_kernel.
__global__ void __launch_bounds__(256)
search_coop4_kernel(const uint2* __restrict__ in, uint2* __restrict__ out,
		const Bucket* __restrict__ table, size_t n, Geom g)
{
	const unsigned lane = threadIdx.x & 31u, sub = lane & 3u, grp0 = lane & ~3u;
	const size_t per_iter = ((size_t)gridDim.x * blockDim.x) >> 2;
	const size_t n_up = (n + 7) & ~(size_t)7;                        // warp-uniform trip count
	for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; i < n_up; i += per_iter) {
		const bool live = i < n;
		uint2 q = make_uint2(0u, 0u);
		if (live) q = ld_stream_u2(in + i);
		const uint32_t b = sub < 2 ? bucket1(g, q.y) : bucket2(g, q.y, q.x);
		uint32_t m = 0, loc = 0;
		if (live) {
			uint4 v = ld_half_row(table[b].sig + 4 * (sub & 1u));
			m = (v.x == q.x ? 1u : 0u) | (v.y == q.x ? 2u : 0u) | (v.z == q.x ? 4u : 0u) | (v.w == q.x ? 8u : 0u);
			if (m) loc = ld_u32_ro(&table[b].loc[4 * (sub & 1u) + (__ffs(m) - 1)]);
		}
		const unsigned hits = (__ballot_sync(0xffffffffu, m != 0) >> grp0) & 0xfu;
		const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 1u) ? 0 : 1));   // lower half wins
		const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 4u) ? 2 : 3));
		if (live && sub == 0)
			st_stream_u2(out + i, make_uint2((hits & 3u) ? l0 : 0u, (hits & 12u) ? l1 : 0u));
	}
}

}  // namespace gh
#endif  /* GH_DEFINE_KERNELS */
