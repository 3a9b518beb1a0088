/*
 * gpuhash_kernels.cuh -- device side of the B200-native Mega-KV hash index.
 *
 * Written from scratch for sm_100a; it is not a port of the reference kernels
 * (pzrq/megakv libgpuhash/gpu_hash.cu).  What it keeps is the *semantics* of that
 * file, cited per function.
 *
 * Work decomposition (DESIGN.md "Kernels"):
 *   reference: 8 lanes per request, one 4 B load per lane, ballot, __syncthreads
 *   here:      what the memory system counts decides the shape (profiles/r01_l2_requests.md: a random probe costs one
 *              128 B line fill and one L2 request per load INSTRUCTION that touches the line).
 *              search: four lanes per request, lane j loads sector j of {b1.lo, b1.hi, b2.lo, b2.hi} with one
 *                      LDG.E.256 -- each bucket is one request; a warp owns a 64-request tile (warp_tile_search).
 *              insert/delete (pair layout): two lanes per request, the bucket again one request, the even lane
 *                      commits with a 64-bit CAS (insert_pair / delete_pair).
 *              One thread per request (a bucket = two 256-bit loads in 16 registers) remains for the reference byte
 *              layout, L2-resident tables and the serial parity kernels.
 *
 * Table layouts (Geom::layout; DESIGN.md "Table layout", profiles/r01_*):
 *   kLayoutPairs  slot l of a bucket is the 8-byte pair {sig, loc} at byte 8*l.  One 64-bit CAS
 *                 publishes, updates, evicts or deletes a (sig, loc) pair atomically, so
 *                 concurrent inserts can never tear a pair.  Default.
 *   kLayoutSplit  the reference's bytes (bucket_t, gpu_hash.h:79-82): 32 B signature row, then
 *                 32 B location row.  For callers that memcpy tables in that layout
 *                 (libgpuhash/test/back/py_search_stream.c:104-121).  Claims/evictions are a CAS
 *                 on the signature word followed by a store/exchange of the location word: the
 *                 same two-step publication the reference has, with its (much narrower) window.
 *   Measured on B200 (ncu, every load flavour): a random 4..32 B read costs a full 128 B line
 *   of HBM traffic, so both layouts cost the same DRAM bytes per probe (2 lines per search);
 *   Pairs just never needs the dependent second load for the location.
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace gh {

constexpr int kSlots = 8;                 // ELEM_NUM          gpu_hash.h:49
constexpr uint32_t kAlgoCuckoo  = 0;      // HASH_CUCKOO       gpu_hash.h:73
constexpr uint32_t kAlgo2Choice = 1;      // HASH_2CHOICE      gpu_hash.h:72
constexpr uint32_t kLayoutPairs = 0;
constexpr uint32_t kLayoutSplit = 1;

struct Geom {                             // == gpuhash_geom_t (gpuhash_ex.h)
	uint32_t hash_mask;                   // buckets of THIS table - 1   (HASH_MASK, gpu_hash.h:61)
	uint32_t block_mask;                  // BLOCK_HASH_MASK of the logical table (gpu_hash.h:69)
	uint32_t algo;
	uint32_t max_cuckoo;                  // MAX_CUCKOO_NUM    gpu_hash.h:75
	uint32_t layout;
};

struct __align__(64) Bucket { uint32_t w[16]; };   // 64 B, stride 64 B in both layouts
struct __align__(32) Row { uint32_t w[kSlots]; };
struct Bkt { Row a, b; };                          // a bucket in registers: words 0..7, 8..15

struct Stats {                            // == gpuhash_stats_t (gpuhash_ex.h)
	unsigned long long ins_skipped, ins_updated, ins_placed_b1, ins_placed_b2, ins_to_b2,
	                   ins_displaced, ins_dropped, ins_overwritten, ins_cas_retry, ins_gave_up,
	                   chain_hist[8],
	                   del_zeroed, del_requests_hit,
	                   search_hits_b1, search_hits_b2;
};

/* ------------------------------------------------------------------ memory ops */

// How much of a missing 128 B line L2 fetches from HBM.  Measured on B200 (tools/gather_flavours, ncu dram__sectors_read per
// random 32 B access, profiles/r02_l2_requests.md): 3.98 sectors with a plain load of any flavour, cache policy or
// cudaLimitMaxL2FetchGranularity -- and 2.00 with the .L2::64B qualifier (SASS LTC64B).  A bucket is 64 B, 64 B aligned: with
// the qualifier a probe moves exactly its bucket and nothing else, halving the DRAM bytes of every table access.  (The RATE of
// random probes does not change: ~47 G/s is a request limit, not a byte limit.)  -DGH_LTC_PLAIN restores plain loads for A/B runs.
#ifdef GH_LTC_PLAIN
#define GH_LTC ""
#else
#define GH_LTC ".L2::64B"
#endif

// Table rows during search.  Random rows have no reuse inside a launch, so they are not
// allocated in L1.  Deliberately NOT .nc: kernels of other streams may be inserting into the
// same table (mega_scheduler.c runs one stream per worker with no cross-stream order), and a
// weak load that returns the old or the new word is what the reference's plain loads do too.
__device__ __forceinline__ Row ld_row_ro(const uint32_t* p)
{
	Row r;
	asm volatile("ld.global.L1::no_allocate" GH_LTC ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
		  "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p));
	return r;
}

// Coherent at L2 (never served from L1): rows read by insert/delete, which race with
// other threads' CAS on the same words.
__device__ __forceinline__ Row ld_row_strong(const uint32_t* p)
{
	Row r;
	asm volatile("ld.relaxed.gpu.global" GH_LTC ".v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]),
		  "=r"(r.w[4]), "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7]) : "l"(p) : "memory");
	return r;
}

__device__ __forceinline__ uint32_t ld_u32_ro(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.global.L1::no_allocate" GH_LTC ".u32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

__device__ __forceinline__ uint32_t ld_u32_strong(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global" GH_LTC ".u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ void st_u32_strong(uint32_t* p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// request stream in / result stream out: touched once -> .cs (streaming, evict-first) so they
// do not displace table lines from L2 when the table is L2-resident
__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p)
{
	uint2 v;
	asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
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
	asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// Programmatic dependent launch: let the next kernel of the stream be scheduled while this one runs, and do not touch
// anything the previous kernel of the stream wrote before it has completed.  Both are no-ops in a launch without the
// programmatic-serialization attribute (libgpuhash.cu: GPUHASH_PDL).
__device__ __forceinline__ void pdl_enter()
{
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	asm volatile("griddepcontrol.wait;" ::: "memory");
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

/* ---- a whole bucket in registers, read through either layout (all indices compile-time) ---- */

template <bool kPairs, int l> __device__ __forceinline__ uint32_t slot_sig(const Bkt& k)
{
	return kPairs ? (l < 4 ? k.a.w[2 * (l & 3)] : k.b.w[2 * (l & 3)]) : k.a.w[l];
}
template <bool kPairs, int l> __device__ __forceinline__ uint32_t slot_loc(const Bkt& k)
{
	return kPairs ? (l < 4 ? k.a.w[2 * (l & 3) + 1] : k.b.w[2 * (l & 3) + 1]) : k.b.w[l];
}

template <bool kPairs> __device__ __forceinline__ uint32_t sig_mask(const Bkt& k, uint32_t v)
{
	return (slot_sig<kPairs, 0>(k) == v ? 1u : 0u)   | (slot_sig<kPairs, 1>(k) == v ? 2u : 0u)
	     | (slot_sig<kPairs, 2>(k) == v ? 4u : 0u)   | (slot_sig<kPairs, 3>(k) == v ? 8u : 0u)
	     | (slot_sig<kPairs, 4>(k) == v ? 16u : 0u)  | (slot_sig<kPairs, 5>(k) == v ? 32u : 0u)
	     | (slot_sig<kPairs, 6>(k) == v ? 64u : 0u)  | (slot_sig<kPairs, 7>(k) == v ? 128u : 0u);
}
template <bool kPairs> __device__ __forceinline__ uint32_t loc_mask(const Bkt& k, uint32_t v)
{
	return (slot_loc<kPairs, 0>(k) == v ? 1u : 0u)   | (slot_loc<kPairs, 1>(k) == v ? 2u : 0u)
	     | (slot_loc<kPairs, 2>(k) == v ? 4u : 0u)   | (slot_loc<kPairs, 3>(k) == v ? 8u : 0u)
	     | (slot_loc<kPairs, 4>(k) == v ? 16u : 0u)  | (slot_loc<kPairs, 5>(k) == v ? 32u : 0u)
	     | (slot_loc<kPairs, 6>(k) == v ? 64u : 0u)  | (slot_loc<kPairs, 7>(k) == v ? 128u : 0u);
}
// sig / loc of slot l for a run-time l: a select chain, no local-memory indexing
template <bool kPairs> __device__ __forceinline__ uint32_t sig_at(const Bkt& k, int l)
{
	uint32_t v = slot_sig<kPairs, 0>(k);
	if (l == 1) v = slot_sig<kPairs, 1>(k);
	if (l == 2) v = slot_sig<kPairs, 2>(k);
	if (l == 3) v = slot_sig<kPairs, 3>(k);
	if (l == 4) v = slot_sig<kPairs, 4>(k);
	if (l == 5) v = slot_sig<kPairs, 5>(k);
	if (l == 6) v = slot_sig<kPairs, 6>(k);
	if (l == 7) v = slot_sig<kPairs, 7>(k);
	return v;
}
template <bool kPairs> __device__ __forceinline__ uint32_t loc_at(const Bkt& k, int l)
{
	uint32_t v = slot_loc<kPairs, 0>(k);
	if (l == 1) v = slot_loc<kPairs, 1>(k);
	if (l == 2) v = slot_loc<kPairs, 2>(k);
	if (l == 3) v = slot_loc<kPairs, 3>(k);
	if (l == 4) v = slot_loc<kPairs, 4>(k);
	if (l == 5) v = slot_loc<kPairs, 5>(k);
	if (l == 6) v = slot_loc<kPairs, 6>(k);
	if (l == 7) v = slot_loc<kPairs, 7>(k);
	return v;
}

__device__ __forceinline__ Bkt ld_bucket_ro(const Bucket* b)
{
	Bkt k; k.a = ld_row_ro(b->w); k.b = ld_row_ro(b->w + 8); return k;
}
__device__ __forceinline__ Bkt ld_bucket_strong(const Bucket* b)
{
	Bkt k; k.a = ld_row_strong(b->w); k.b = ld_row_strong(b->w + 8); return k;
}

/* ------------------------------------------------------------------ search */

// gpu_hash.cu:47-72 for one request.  Both buckets are always probed (the early exit at :61-63
// is commented out in the reference).  With several matching slots the LOWEST one is reported:
// the reference lets all matching lanes store to the same word, and run on a B200 its kernel
// keeps the lowest lane's store (tests/golden/ref_search_cuckoo_16.npz, dup_* arrays).
//
// kSearchPairs       Pairs layout, whole buckets (4 x LDG.256, no dependent load)
// kSearchSplitWhole  Split layout, whole buckets (signature AND location row up front; same DRAM lines)
// kSearchSplitLazy   Split layout, signature rows first, location word only on a hit (fewest L2
//                    sectors: chosen when the table fits in L2, where sectors -- not DRAM lines --
//                    are the cost)
constexpr int kSearchPairs = 0, kSearchSplitWhole = 1, kSearchSplitLazy = 2;

template <int kMode> struct Probe { uint32_t b1, b2; Bkt k1, k2; };
template <> struct Probe<kSearchSplitLazy> { uint32_t b1, b2; Row r1, r2; };

template <int kMode>
__device__ __forceinline__ void search_issue(const Bucket* __restrict__ table, const Geom& g,
		uint2 q /* x = sig, y = hash */, Probe<kMode>& p)
{
	p.b1 = bucket1(g, q.y);
	p.b2 = bucket2(g, q.y, q.x);
	if constexpr (kMode == kSearchSplitLazy) {
		p.r1 = ld_row_ro(table[p.b1].w); p.r2 = ld_row_ro(table[p.b2].w);
	} else {
		p.k1 = ld_bucket_ro(table + p.b1); p.k2 = ld_bucket_ro(table + p.b2);
	}
}

template <int kMode>
__device__ __forceinline__ uint2 search_finish(const Bucket* __restrict__ table, uint2 q,
		const Probe<kMode>& p, uint32_t& hit1, uint32_t& hit2)
{
	if constexpr (kMode == kSearchSplitLazy) {
		hit1 = eq_mask(p.r1, q.x); hit2 = eq_mask(p.r2, q.x);
		// the two location loads are independent: issue both before either is consumed
		const uint32_t* p1 = &table[p.b1].w[8 + ((__ffs(hit1 | 0x100u) - 1) & 7)];
		const uint32_t* p2 = &table[p.b2].w[8 + ((__ffs(hit2 | 0x100u) - 1) & 7)];
		uint32_t v1 = 0, v2 = 0;
		if (hit1) v1 = ld_u32_ro(p1);
		if (hit2) v2 = ld_u32_ro(p2);
		return make_uint2(v1, v2);
	} else {
		constexpr bool kPairs = kMode == kSearchPairs;
		hit1 = sig_mask<kPairs>(p.k1, q.x); hit2 = sig_mask<kPairs>(p.k2, q.x);
		uint32_t v1 = hit1 ? loc_at<kPairs>(p.k1, __ffs(hit1) - 1) : 0u;
		uint32_t v2 = hit2 ? loc_at<kPairs>(p.k2, __ffs(hit2) - 1) : 0u;
		return make_uint2(v1, v2);
	}
}

// run-time layout dispatch for callers outside libgpuhash.cu (the sharded lookup kernel)
__device__ __forceinline__ uint2 search_one(const Bucket* __restrict__ table, const Geom& g, uint2 q)
{
	uint32_t h1, h2;
	if (g.layout == kLayoutPairs) {
		Probe<kSearchPairs> p; search_issue<kSearchPairs>(table, g, q, p);
		return search_finish<kSearchPairs>(table, q, p, h1, h2);
	}
	Probe<kSearchSplitWhole> p; search_issue<kSearchSplitWhole>(table, g, q, p);
	return search_finish<kSearchSplitWhole>(table, q, p, h1, h2);
}

#ifdef GH_DEFINE_KERNELS   /* __global__ definitions: only libgpuhash.cu instantiates them */
// One thread per request, kQpt requests per thread issued back to back so that each thread
// keeps 2*kQpt (lazy) or 4*kQpt (whole) 32 B reads in flight.  `out` gets both words of every
// request (0 = miss): the caller's cudaMemset of `out` (mega_scheduler.c:406) is fused away.
template <int kQpt, int kMode>
__global__ void __launch_bounds__(256)
search_kernel(const uint2* __restrict__ in, uint2* __restrict__ out,
		const Bucket* __restrict__ table, size_t n, Geom g, Stats* st)
{
	const size_t tile = (size_t)blockDim.x * kQpt;
	for (size_t base = (size_t)blockIdx.x * tile; base < n; base += (size_t)gridDim.x * tile) {
		uint2 q[kQpt]; Probe<kMode> p[kQpt];
		bool live[kQpt];
#pragma unroll
		for (int k = 0; k < kQpt; k++) {
			size_t i = base + (size_t)k * blockDim.x + threadIdx.x;
			live[k] = i < n;
			if (live[k]) q[k] = ld_stream_u2(in + i);
		}
#pragma unroll
		for (int k = 0; k < kQpt; k++)
			if (live[k]) search_issue<kMode>(table, g, q[k], p[k]);
#pragma unroll
		for (int k = 0; k < kQpt; k++) {
			if (!live[k]) continue;
			size_t i = base + (size_t)k * blockDim.x + threadIdx.x;
			uint32_t h1, h2;
			uint2 o = search_finish<kMode>(table, q[k], p[k], h1, h2);
			st_stream_u2(out + i, o);
			if (st) {
				if (h1) atomicAdd(&st->search_hits_b1, 1ULL);
				if (h2) atomicAdd(&st->search_hits_b2, 1ULL);
			}
		}
	}
}
#endif  /* GH_DEFINE_KERNELS */

/* ------------------------------------------------------------------ delete */

// gpu_hash.cu:454-477 for one request: zero the signature of every slot whose signature AND
// location match (the location word stays); visit bucket 2 only if this request zeroed nothing
// in bucket 1 (:465-468).  The zeroing is a CAS, so two identical requests of one batch behave
// like the sequential run: the first zeroes, the second sees a miss and goes on to bucket 2.
template <bool kPairs>
__device__ __forceinline__ int delete_in_bucket(Bucket* bk, uint32_t sig, uint32_t loc)
{
	Bkt k = ld_bucket_strong(bk);
	uint32_t m = sig_mask<kPairs>(k, sig) & loc_mask<kPairs>(k, loc);
	int zeroed = 0;
	while (m) {
		int l = __ffs(m) - 1; m &= m - 1;
		if (kPairs) {
			unsigned long long expect = ((unsigned long long)loc << 32) | sig;
			unsigned long long* slot = (unsigned long long*)&bk->w[2 * l];
			if (atomicCAS(slot, expect, (unsigned long long)loc << 32) == expect) zeroed++;
		} else {
			if (atomicCAS(&bk->w[l], sig, 0u) == sig) zeroed++;
		}
	}
	return zeroed;
}

template <bool kPairs>
__device__ __forceinline__ int delete_one(Bucket* table, const Geom& g,
		uint32_t sig, uint32_t hash, uint32_t loc)
{
	int z = delete_in_bucket<kPairs>(table + bucket1(g, hash), sig, loc);
	if (z) return z;
	return delete_in_bucket<kPairs>(table + bucket2(g, hash, sig), sig, loc);
}

/* ------------------------------------------------------------------ insert */

#define GH_COUNT(field) do { if (st) atomicAdd(&st->field, 1ULL); } while (0)

// gpu_hash.cu:256-430 (cuckoo) and :97-226 (2-choice) for one request, as a bounded lock-free
// loop.  One iteration = "read the current bucket, decide, commit with one CAS":
//   signature present  -> new location (update in place)              :277-287, 339-349
//   empty slot         -> claim it; slot = first empty from the major location sig0 & 7
//                                                                      :303-327, 351-395
//   bucket 1 full      -> go to the alternate bucket                   :330-336
//   alternate full     -> cuckoo:  victim slot sig0 & 7 is replaced and the victim carried on --
//                                  with the REQUEST's hash, never the victim's (:334-335,
//                                  403-404) -- at most max_cuckoo times, then the victim is
//                                  overwritten without re-homing (:414-422)
//                         2-choice: signature stored into slot sig & 7, location untouched (:197-209)
// Pairs layout: every commit is ONE 64-bit CAS on the {sig, loc} slot, expected value = what
// this thread read.  A failed CAS means another request changed that slot first: the bucket is
// read again and the decision retaken -- the order "other request, then this one" of a
// sequential run.  The table therefore always equals SOME sequential execution of the batch.
// Split layout: the CAS is on the signature word; the location follows with a store (claim,
// update) or an exchange (eviction).  Two requests that take the same slot within that gap can
// pair a signature with the other's location -- the reference's own publication window.
// Every failed CAS is another request's success, so the system as a whole always advances
// (lock-free); kMaxSteps additionally bounds one request's own loop (counted in ins_gave_up,
// asserted 0 by the tests even with 40 000 requests aimed at 64 buckets).
constexpr int kMaxSteps = 1 << 16;

template <bool kPairs>
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
		Bkt k;
		if (kPairs) k = ld_bucket_strong(bk);
		else { k.a = ld_row_strong(bk->w); k.b = k.a; }              // Split: the signature row is enough
		const uint32_t hit = sig_mask<kPairs>(k, sig);
		if (hit) {                                                   // update in place, lowest slot
			const int l = __ffs(hit) - 1;
			if (kPairs) {
				unsigned long long expect = ((unsigned long long)loc_at<true>(k, l) << 32) | sig;
				unsigned long long want = ((unsigned long long)loc << 32) | sig;
				if (expect != want && atomicCAS((unsigned long long*)&bk->w[2 * l], expect, want) != expect) {
					GH_COUNT(ins_cas_retry); continue;
				}
			} else if (ld_u32_strong(&bk->w[8 + l]) != loc) {         // (a hot key updated with the location it already has -- zipf SETs --
				st_u32_strong(&bk->w[8 + l], loc);                    //  would otherwise serialise thousands of stores on one word)
			}
			GH_COUNT(ins_updated);
			goto done;
		}
		const uint32_t empty = sig_mask<kPairs>(k, 0u);
		if (empty) {
			const int l = first_from(empty, major);
			if (kPairs) {
				unsigned long long expect = (unsigned long long)loc_at<true>(k, l) << 32;   // {0, stale loc}
				unsigned long long want = ((unsigned long long)loc << 32) | sig;
				if (atomicCAS((unsigned long long*)&bk->w[2 * l], expect, want) != expect) {
					GH_COUNT(ins_cas_retry); continue;
				}
			} else {
				const uint32_t old = atomicCAS(&bk->w[l], 0u, sig);
				if (old != 0u && old != sig) { GH_COUNT(ins_cas_retry); continue; }
				st_u32_strong(&bk->w[8 + l], loc);
				if (old == sig) { GH_COUNT(ins_updated); goto done; }   // a twin claimed it first
			}
			if (alt) GH_COUNT(ins_placed_b2); else GH_COUNT(ins_placed_b1);
			goto done;
		}
		if (!alt) {                                                  // bucket 1 full
			alt = true;
			GH_COUNT(ins_to_b2);
			b = bucket2(g, hash, sig);
			continue;
		}
		// alternate bucket full
		const int l = (int)(sig0 & (kSlots - 1));                    // elem->sig :200, 360
		const uint32_t vsig = sig_at<kPairs>(k, l);
		if (g.algo == kAlgo2Choice) {                                // signature only :197-209
			if (kPairs) {
				const unsigned long long vloc = loc_at<true>(k, l);
				unsigned long long expect = (vloc << 32) | vsig, want = (vloc << 32) | sig;
				if (atomicCAS((unsigned long long*)&bk->w[2 * l], expect, want) != expect) {
					GH_COUNT(ins_cas_retry); continue;
				}
			} else {
				st_u32_strong(&bk->w[l], sig);
			}
			GH_COUNT(ins_overwritten);
			goto done;
		}
		uint32_t vloc;
		if (kPairs) {
			vloc = loc_at<true>(k, l);
			unsigned long long expect = ((unsigned long long)vloc << 32) | vsig;
			unsigned long long want = ((unsigned long long)loc << 32) | sig;
			if (atomicCAS((unsigned long long*)&bk->w[2 * l], expect, want) != expect) {
				GH_COUNT(ins_cas_retry); continue;
			}
		} else {
			if (atomicCAS(&bk->w[l], vsig, sig) != vsig) { GH_COUNT(ins_cas_retry); continue; }
			vloc = atomicExch(&bk->w[8 + l], loc);
		}
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

/* ---- two lanes per request (pair layout): the whole 64 B bucket is ONE L2 request ---- */

// insert_one / delete_one read a bucket with two load instructions of one thread: two L2 requests for one 128 B line
// (profiles/r01_l2_requests.md).  Here lanes 2p and 2p+1 serve one request: each loads one half of the bucket in the SAME
// instruction (one request), the halves are exchanged by shuffles so that both lanes hold the bucket and take the same
// decision, the even lane commits it with the 64-bit CAS and tells its partner how that went.  Same decisions, same
// CAS discipline, same counters as insert_one / delete_one; only the loads differ.  All 32 lanes run the loop until the
// last pair of the warp is done (the shuffles are warp-wide).
__device__ __forceinline__ Bkt ld_bucket_pair(const Bucket* bk, bool active, unsigned h)
{
	Row mine, other;
#pragma unroll
	for (int w = 0; w < 8; w++) mine.w[w] = 0;
	if (active) mine = ld_row_strong(bk->w + 8 * h);
#pragma unroll
	for (int w = 0; w < 8; w++) other.w[w] = __shfl_xor_sync(0xffffffffu, mine.w[w], 1);
	Bkt k;
	k.a = h ? other : mine; k.b = h ? mine : other;
	return k;
}

__device__ __forceinline__ void insert_pair(Bucket* table, const Geom& g, bool have,
		uint32_t sig0, uint32_t hash, uint32_t loc0, Stats* st, unsigned lane)
{
	const unsigned h = lane & 1u;
	const bool count = st && h == 0;
	bool active = have && !(sig0 == 0 && loc0 == 0);
	if (have && !active && count) atomicAdd(&st->ins_skipped, 1ULL);        // :101-104, 259-262
	uint32_t sig = sig0, loc = loc0;
	const int major = (int)(sig0 & (kSlots - 1));
	uint32_t b = bucket1(g, hash);
	bool alt = false;
	uint32_t c = 0;
	int step = 0;
	enum { kUpdate, kClaim, kGoAlt, kOverwrite, kEvict };
	while (__any_sync(0xffffffffu, active)) {
		Bucket* bk = table + b;
		const Bkt k = ld_bucket_pair(bk, active, h);
		int kind = kGoAlt, l = 0;
		unsigned long long expect = 0, want = 0;
		uint32_t vsig = 0, vloc = 0;
		if (active) {
			const uint32_t hit = sig_mask<true>(k, sig);
			const uint32_t empty = sig_mask<true>(k, 0u);
			if (hit) {                                                       // update in place, lowest slot
				kind = kUpdate; l = __ffs(hit) - 1;
				expect = ((unsigned long long)loc_at<true>(k, l) << 32) | sig;
				want = ((unsigned long long)loc << 32) | sig;
			} else if (empty) {                                              // claim, first empty from the major location
				kind = kClaim; l = first_from(empty, major);
				expect = (unsigned long long)loc_at<true>(k, l) << 32;
				want = ((unsigned long long)loc << 32) | sig;
			} else if (!alt) {
				kind = kGoAlt;                                               // bucket 1 full
			} else {                                                         // alternate bucket full
				l = major; vsig = sig_at<true>(k, l); vloc = loc_at<true>(k, l);
				expect = ((unsigned long long)vloc << 32) | vsig;
				if (g.algo == kAlgo2Choice) { kind = kOverwrite; want = ((unsigned long long)vloc << 32) | sig; }
				else                        { kind = kEvict;     want = ((unsigned long long)loc << 32) | sig; }
			}
		}
		int ok = 1;
		if (active && h == 0 && kind != kGoAlt && expect != want)
			ok = atomicCAS((unsigned long long*)&bk->w[2 * l], expect, want) == expect;
		ok = __shfl_sync(0xffffffffu, ok, lane & ~1u);
		if (active) {
			if (kind == kGoAlt) {
				alt = true; b = bucket2(g, hash, sig);
				if (count) atomicAdd(&st->ins_to_b2, 1ULL);
			} else if (!ok) {
				if (count) atomicAdd(&st->ins_cas_retry, 1ULL);
			} else if (kind == kUpdate) {
				if (count) atomicAdd(&st->ins_updated, 1ULL);
				active = false;
			} else if (kind == kClaim) {
				if (count) atomicAdd(alt ? &st->ins_placed_b2 : &st->ins_placed_b1, 1ULL);
				active = false;
			} else if (kind == kOverwrite) {
				if (count) atomicAdd(&st->ins_overwritten, 1ULL);
				active = false;
			} else {                                                         // kEvict
				if (c < g.max_cuckoo) {
					c++;
					if (count) atomicAdd(&st->ins_displaced, 1ULL);
					sig = vsig; loc = vloc;
					b = bucket2(g, hash, sig);                               // request's hash, victim's sig (:334-335)
				} else {
					if (count) atomicAdd(&st->ins_dropped, 1ULL);
					active = false;
				}
			}
			if (active && ++step >= kMaxSteps) { if (count) atomicAdd(&st->ins_gave_up, 1ULL); active = false; }
			if (!active && count && g.algo == kAlgoCuckoo) atomicAdd(&st->chain_hist[c < 7 ? c : 7], 1ULL);
		}
	}
}

__device__ __forceinline__ int delete_pair(Bucket* table, const Geom& g, bool have,
		uint32_t sig, uint32_t hash, uint32_t loc, unsigned lane)
{
	const unsigned h = lane & 1u;
	int zeroed = 0;
#pragma unroll
	for (int round = 0; round < 2; round++) {
		const bool active = have && zeroed == 0;                             // bucket 2 only if nothing was zeroed in bucket 1 (:465-468)
		Bucket* bk = table + (round == 0 ? bucket1(g, hash) : bucket2(g, hash, sig));
		const Bkt k = ld_bucket_pair(bk, active, h);
		int z = 0;
		if (active && h == 0) {
			uint32_t m = sig_mask<true>(k, sig) & loc_mask<true>(k, loc);
			const unsigned long long expect = ((unsigned long long)loc << 32) | sig;
			while (m) {
				const int l = __ffs(m) - 1; m &= m - 1;
				if (atomicCAS((unsigned long long*)&bk->w[2 * l], expect, (unsigned long long)loc << 32) == expect) z++;
			}
		}
		z = __shfl_sync(0xffffffffu, z, lane & ~1u);
		if (active) zeroed = z;
	}
	return zeroed;
}

#ifdef GH_DEFINE_KERNELS
template <bool kPairs>
__global__ void __launch_bounds__(256)
delete_kernel(const uint32_t* __restrict__ in /* delem_t[n] as words */, Bucket* table,
		size_t n, Geom g, Stats* st)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
			i += (size_t)gridDim.x * blockDim.x) {
		uint32_t sig = ld_stream_u32(in + 3 * i), hash = ld_stream_u32(in + 3 * i + 1),
		         loc = ld_stream_u32(in + 3 * i + 2);
		int z = delete_one<kPairs>(table, g, sig, hash, loc);
		if (st && z) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
	}
}

// The legacy entry point only knows num_blks on the host; segment sizes live in device
// memory (mega_scheduler.c:493-494).  So the grid is count-independent: every CTA builds
// the prefix sum of the segment sizes in shared memory and the grid strides over the
// concatenation.  Segment membership carries no meaning.
constexpr int kMaxSegChunk = 1024;

template <bool kPairs>
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
		if (kPairs) {                                                // two lanes per request, warp-uniform trip count
			const unsigned long long total_up = (total + 15ULL) & ~15ULL;
			for (unsigned long long e = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
					e < total_up; e += ((unsigned long long)gridDim.x * blockDim.x) >> 1) {
				const bool have = e < total;
				uint32_t a = 0, b = 0, c = 0;
				if (have) {
					while (e >= prefix[k + 1]) k++;                  // e only grows
					const uint32_t* p = base[k] + 3 * (e - prefix[k]);
					a = ld_stream_u32(p); b = ld_stream_u32(p + 1); c = ld_stream_u32(p + 2);
				}
				insert_pair(table, g, have, a, b, c, st, threadIdx.x & 31u);
			}
		} else {
			for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
					e < total; e += (unsigned long long)gridDim.x * blockDim.x) {
				while (e >= prefix[k + 1]) k++;                      // e only grows
				const uint32_t* p = base[k] + 3 * (e - prefix[k]);
				insert_one<kPairs>(table, g, ld_stream_u32(p), ld_stream_u32(p + 1), ld_stream_u32(p + 2), st);
			}
		}
	}
}

// two lanes per request (pair layout only): warp-uniform trip count, lanes beyond n idle but take part in the shuffles
__global__ void __launch_bounds__(256)
insert_flat_pair_kernel(Bucket* table, const uint32_t* __restrict__ in, size_t n, Geom g, Stats* st)
{
	pdl_enter();
	const unsigned lane = threadIdx.x & 31u;
	const size_t per_iter = ((size_t)gridDim.x * blockDim.x) >> 1;
	const size_t n_up = (n + 15) & ~(size_t)15;
	for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1; i < n_up; i += per_iter) {
		const bool have = i < n;
		uint32_t sig = 0, hash = 0, loc = 0;
		if (have) { sig = ld_stream_u32(in + 3 * i); hash = ld_stream_u32(in + 3 * i + 1); loc = ld_stream_u32(in + 3 * i + 2); }
		insert_pair(table, g, have, sig, hash, loc, st, lane);
	}
}

__global__ void __launch_bounds__(256)
delete_pair_kernel(const uint32_t* __restrict__ in, Bucket* table, size_t n, Geom g, Stats* st)
{
	const unsigned lane = threadIdx.x & 31u;
	const size_t per_iter = ((size_t)gridDim.x * blockDim.x) >> 1;
	const size_t n_up = (n + 15) & ~(size_t)15;
	for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 1; i < n_up; i += per_iter) {
		const bool have = i < n;
		uint32_t sig = 0, hash = 0, loc = 0;
		if (have) { sig = ld_stream_u32(in + 3 * i); hash = ld_stream_u32(in + 3 * i + 1); loc = ld_stream_u32(in + 3 * i + 2); }
		const int z = delete_pair(table, g, have, sig, hash, loc, lane);
		if (st && z && (lane & 1u) == 0) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
	}
}

// host-known count (extended API, pipeline)
template <bool kPairs>
__global__ void __launch_bounds__(256)
insert_flat_kernel(Bucket* table, const uint32_t* __restrict__ in, size_t n, Geom g, Stats* st)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
			i += (size_t)gridDim.x * blockDim.x)
		insert_one<kPairs>(table, g, ld_stream_u32(in + 3 * i), ld_stream_u32(in + 3 * i + 1),
				ld_stream_u32(in + 3 * i + 2), st);
}

// Sequential execution of the SAME device code by one thread, segments in order, requests in
// order: reproduces the oracle slot for slot at any load factor.  Used by the parity tests
// (GPUHASH_INSERT_SERIAL); not a fast path.
template <bool kPairs>
__global__ void insert_serial_kernel(Bucket* table, const uint32_t* const* blk_input,
		const int* blk_elem_num, int num_blks, const uint32_t* flat, size_t flat_n, Geom g, Stats* st)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	if (flat) {
		for (size_t i = 0; i < flat_n; i++)
			insert_one<kPairs>(table, g, flat[3 * i], flat[3 * i + 1], flat[3 * i + 2], st);
		return;
	}
	for (int k = 0; k < num_blks; k++) {
		const uint32_t* p = blk_input[k];
		for (int i = 0; i < blk_elem_num[k]; i++)
			insert_one<kPairs>(table, g, p[3 * i], p[3 * i + 1], p[3 * i + 2], st);
	}
}

template <bool kPairs>
__global__ void delete_serial_kernel(const uint32_t* in, Bucket* table, size_t n, Geom g, Stats* st)
{
	if (blockIdx.x != 0 || threadIdx.x != 0) return;
	for (size_t i = 0; i < n; i++) {
		int z = delete_one<kPairs>(table, g, in[3 * i], in[3 * i + 1], in[3 * i + 2]);
		if (st && z) { atomicAdd(&st->del_zeroed, (unsigned long long)z); atomicAdd(&st->del_requests_hit, 1ULL); }
	}
}

// In-place conversion between the two layouts (one thread per bucket):
//   to_pairs: words (l, 8+l) -> (2l, 2l+1);  !to_pairs: the inverse.
__global__ void __launch_bounds__(256)
convert_layout_kernel(Bucket* table, size_t buckets, int to_pairs)
{
	for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < buckets;
			b += (size_t)gridDim.x * blockDim.x) {
		Bkt k = ld_bucket_strong(table + b);
		uint32_t in[16], out[16];
#pragma unroll
		for (int i = 0; i < 8; i++) { in[i] = k.a.w[i]; in[8 + i] = k.b.w[i]; }
#pragma unroll
		for (int l = 0; l < 8; l++) {
			if (to_pairs) { out[2 * l] = in[l]; out[2 * l + 1] = in[8 + l]; }
			else          { out[l] = in[2 * l]; out[8 + l] = in[2 * l + 1]; }
		}
		uint4* dst = (uint4*)table[b].w;
#pragma unroll
		for (int i = 0; i < 4; i++) dst[i] = make_uint4(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
	}
}

/* ---- four lanes per request: every bucket is ONE L2 request ---- */

// Lane j of a 4-lane group loads sector j of the request's four sectors {b1.lo, b1.hi, b2.lo, b2.hi} in the SAME
// load instruction, so the two sectors of a bucket leave the SM as one request for one 128 B line.  Measured on
// B200 (tools/gather_flavours, profiles/): what is scarce beyond L2 is requests to lines that are not resident
// (~47 G/s), not bytes; two sectors asked for by two instructions cost two requests, asked for by two lanes of one
// instruction they cost one.  Results identical to search_kernel.
template <bool kPairs>
__global__ void __launch_bounds__(256)
search_quad_kernel(const uint2* __restrict__ in, uint2* __restrict__ out,
		const Bucket* __restrict__ table, size_t n, Geom g, Stats* st)
{
	const unsigned lane = threadIdx.x & 31u, sub = lane & 3u, grp0 = lane & ~3u, half = sub & 1u;
	const size_t per_iter = ((size_t)gridDim.x * blockDim.x) >> 2;
	const size_t n_up = (n + 7) & ~(size_t)7;                        // warp-uniform trip count
	for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; i < n_up; i += per_iter) {
		const bool live = i < n;
		uint2 q = make_uint2(0u, 0u);
		Row r;
#pragma unroll
		for (int k = 0; k < 8; k++) r.w[k] = 0;
		if (live) {
			q = ld_stream_u2(in + i);
			const uint32_t b = sub < 2 ? bucket1(g, q.y) : bucket2(g, q.y, q.x);
			r = ld_row_ro(table[b].w + 8 * half);
		}
		uint32_t m, loc = 0;
		if (kPairs) {                                                // this lane holds slots 4*half .. 4*half+3 as {sig, loc}
			m = (r.w[0] == q.x ? 1u : 0u) | (r.w[2] == q.x ? 2u : 0u) | (r.w[4] == q.x ? 4u : 0u) | (r.w[6] == q.x ? 8u : 0u);
			loc = (m & 1u) ? r.w[1] : (m & 2u) ? r.w[3] : (m & 4u) ? r.w[5] : r.w[7];
			if (!live) m = 0;
			const unsigned hits = (__ballot_sync(0xffffffffu, m != 0) >> grp0) & 0xfu;
			const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 1u) ? 0 : 1));   // lowest slot wins
			const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 4u) ? 2 : 3));
			if (live && sub == 0) {
				st_stream_u2(out + i, make_uint2((hits & 3u) ? l0 : 0u, (hits & 12u) ? l1 : 0u));
				if (st) { if (hits & 3u) atomicAdd(&st->search_hits_b1, 1ULL); if (hits & 12u) atomicAdd(&st->search_hits_b2, 1ULL); }
			}
		} else {                                                     // even lanes hold a signature row, odd lanes its location row
			m = live ? eq_mask(r, q.x) : 0u;
			const uint32_t msig = __shfl_sync(0xffffffffu, m, grp0 + (sub & 2u));            // mask of this bucket's signature lane
			const int l = __ffs(msig | 0x100u) - 1 & 7;
			loc = r.w[0];
			if (l == 1) loc = r.w[1];
			if (l == 2) loc = r.w[2];
			if (l == 3) loc = r.w[3];
			if (l == 4) loc = r.w[4];
			if (l == 5) loc = r.w[5];
			if (l == 6) loc = r.w[6];
			if (l == 7) loc = r.w[7];
			if (!msig) loc = 0;
			const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + 1);
			const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + 3);
			const uint32_t m2 = __shfl_sync(0xffffffffu, m, grp0 + 2);
			if (live && sub == 0) {
				st_stream_u2(out + i, make_uint2(l0, l1));
				if (st) { if (m) atomicAdd(&st->search_hits_b1, 1ULL); if (m2) atomicAdd(&st->search_hits_b2, 1ULL); }
			}
		}
	}
}

#endif  /* GH_DEFINE_KERNELS */

/* ---- four lanes per request + the batch staged through shared memory by bulk copies (TMA 1-D) ---- */

// The request batch and the result batch are moved 64 requests (512 B) at a time by cp.async.bulk: one elected
// thread asks the copy unit for the tile's queries (global -> shared, completion on an mbarrier), every 4-lane
// group reads its query from shared memory, and the 64 results leave as one 512 B bulk store (shared -> global).
// Why: (1) the batches may live in the caller's PINNED HOST buffers (gpuhash_index_set_zero_copy) -- over PCIe a
// warp's 64 B query load / 64 B result store are 64 B transactions and the link tops out at ~21 GB/s per direction
// (tools/pcie_probe), 512 B transfers keep it at the 40+ GB/s large copies reach; (2) in HBM it halves the L2
// requests spent on the streams (0.25 -> 0.125 per search) and takes both off the LSU path.  Two stages: the
// queries of the CTA's next tile are in flight while this one is probed; the store of tile t drains while t+1 runs.
// Needs 16 B-aligned tiles: `head` (0 or 1) leading requests are done the plain way so that in + head is aligned;
// out_bulk says whether out + head is too (else results are stored per request).  The last, partial tile is plain.
constexpr int kTileReq = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}

template <bool kPairs>
__device__ __forceinline__ uint2 quad_probe(const Bucket* __restrict__ table, const Geom& g, uint2 q, bool live,
		unsigned sub, unsigned grp0, unsigned half, uint32_t& hit1, uint32_t& hit2)
{
	Row r;
#pragma unroll
	for (int k = 0; k < 8; k++) r.w[k] = 0;
	if (live) {
		const uint32_t b = sub < 2 ? bucket1(g, q.y) : bucket2(g, q.y, q.x);
		r = ld_row_ro(table[b].w + 8 * half);
	}
	if (kPairs) {                                                // this lane holds slots 4*half .. 4*half+3 as {sig, loc}
		uint32_t m = (r.w[0] == q.x ? 1u : 0u) | (r.w[2] == q.x ? 2u : 0u) | (r.w[4] == q.x ? 4u : 0u) | (r.w[6] == q.x ? 8u : 0u);
		const uint32_t loc = (m & 1u) ? r.w[1] : (m & 2u) ? r.w[3] : (m & 4u) ? r.w[5] : r.w[7];
		if (!live) m = 0;
		const unsigned hits = (__ballot_sync(0xffffffffu, m != 0) >> grp0) & 0xfu;
		const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 1u) ? 0 : 1));   // lowest slot wins
		const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 4u) ? 2 : 3));
		hit1 = hits & 3u; hit2 = hits & 12u;
		return make_uint2(hit1 ? l0 : 0u, hit2 ? l1 : 0u);
	} else {                                                     // even lanes hold a signature row, odd lanes its location row
		const uint32_t m = live ? eq_mask(r, q.x) : 0u;
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
		const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + 1);
		const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + 3);
		hit1 = __shfl_sync(0xffffffffu, m, grp0);
		hit2 = __shfl_sync(0xffffffffu, m, grp0 + 2);
		return make_uint2(l0, l1);
	}
}

/* ---- a warp on its own: 64 requests in, 64 results out, no shared memory, no barrier ---- */

// The warp reads its tile of 64 requests as ONE 512 B access (16 B per lane), hands the requests to its eight 4-lane
// groups by shuffles, probes them in two blocks of four rounds (so every lane has four independent 32 B table loads in
// flight), routes the results back so that lane l holds results 2l and 2l+1, and writes them as ONE 512 B access.
// kSys: the batches are in pinned host memory read/written by a persistent kernel -> system-scope accesses (a weak
// load could be served from a stale L2 line of a reused host buffer); else streaming (.cs) accesses.
// `valid` (1..64) requests exist; in_t / out_t are 16 B aligned (out_vec false: results stored per request).
// kCompact: one word per request instead of two -- the first non-zero of {bucket-1 hit, bucket-2 hit}, which is what the
// consumer takes anyway (src/mega_send.c:411-414); out_t is then uint32_t[64] (8 B aligned for the vector store).
template <bool kPairs, bool kSys, bool kCompact = false>
__device__ __forceinline__ void warp_tile_search(const Bucket* __restrict__ table, const Geom& g,
		const uint2* in_t, void* out_t_, uint32_t valid, bool out_vec, uint4 v /* this lane's two requests, already loaded */,
		unsigned lane, uint32_t& hits1, uint32_t& hits2)
{
	const unsigned sub = lane & 3u, grp0 = lane & ~3u, half = sub & 1u, j = lane >> 2;
	uint4 res = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
	for (int blk = 0; blk < 2; blk++) {
		uint2 q[4]; bool live[4]; Row r[4];
#pragma unroll
		for (int kk = 0; kk < 4; kk++) {                         // round k: group j takes request 8k + j = lane 4k + j/2, half j&1
			const int k = 4 * blk + kk;
			const int src = 4 * k + (int)(j >> 1);
			const uint32_t a = __shfl_sync(0xffffffffu, v.x, src), b = __shfl_sync(0xffffffffu, v.y, src);
			const uint32_t c = __shfl_sync(0xffffffffu, v.z, src), d = __shfl_sync(0xffffffffu, v.w, src);
			q[kk] = (j & 1u) ? make_uint2(c, d) : make_uint2(a, b);
			live[kk] = (uint32_t)(8 * k) + j < valid;
#pragma unroll
			for (int w = 0; w < 8; w++) r[kk].w[w] = 0;
			if (live[kk]) {
				const uint32_t bk = sub < 2 ? bucket1(g, q[kk].y) : bucket2(g, q[kk].y, q[kk].x);
				r[kk] = ld_row_ro(table[bk].w + 8 * half);
			}
		}
#pragma unroll
		for (int kk = 0; kk < 4; kk++) {
			const int k = 4 * blk + kk;
			uint2 o;
			if (kPairs) {
				uint32_t m = (r[kk].w[0] == q[kk].x ? 1u : 0u) | (r[kk].w[2] == q[kk].x ? 2u : 0u) | (r[kk].w[4] == q[kk].x ? 4u : 0u) | (r[kk].w[6] == q[kk].x ? 8u : 0u);
				const uint32_t loc = (m & 1u) ? r[kk].w[1] : (m & 2u) ? r[kk].w[3] : (m & 4u) ? r[kk].w[5] : r[kk].w[7];
				if (!live[kk]) m = 0;
				const unsigned hits = (__ballot_sync(0xffffffffu, m != 0) >> grp0) & 0xfu;
				const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 1u) ? 0 : 1));   // lowest slot wins
				const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 4u) ? 2 : 3));
				o = make_uint2((hits & 3u) ? l0 : 0u, (hits & 12u) ? l1 : 0u);
				if (sub == 0) { hits1 += (hits & 3u) ? 1u : 0u; hits2 += (hits & 12u) ? 1u : 0u; }
			} else {
				const uint32_t m = live[kk] ? eq_mask(r[kk], q[kk].x) : 0u;
				const uint32_t msig = __shfl_sync(0xffffffffu, m, grp0 + (sub & 2u));
				const int l = __ffs(msig | 0x100u) - 1 & 7;
				uint32_t loc = r[kk].w[0];
				if (l == 1) loc = r[kk].w[1];
				if (l == 2) loc = r[kk].w[2];
				if (l == 3) loc = r[kk].w[3];
				if (l == 4) loc = r[kk].w[4];
				if (l == 5) loc = r[kk].w[5];
				if (l == 6) loc = r[kk].w[6];
				if (l == 7) loc = r[kk].w[7];
				if (!msig) loc = 0;
				o = make_uint2(__shfl_sync(0xffffffffu, loc, grp0 + 1), __shfl_sync(0xffffffffu, loc, grp0 + 3));
				const uint32_t m2 = __shfl_sync(0xffffffffu, m, grp0 + 2);
				if (sub == 0) { hits1 += m ? 1u : 0u; hits2 += m2 ? 1u : 0u; }
			}
			// results of round k belong to lanes 4k .. 4k+3: lane 4k+m takes groups 2m (first half) and 2m+1 (second half)
			const int s0 = 8 * (int)(lane & 3u), s1 = s0 + 4;
			const uint32_t rx = __shfl_sync(0xffffffffu, o.x, s0), ry = __shfl_sync(0xffffffffu, o.y, s0);
			const uint32_t rz = __shfl_sync(0xffffffffu, o.x, s1), rw = __shfl_sync(0xffffffffu, o.y, s1);
			if ((int)(lane >> 2) == k) res = make_uint4(rx, ry, rz, rw);
		}
	}
	const uint32_t first = 2 * lane;                             // this lane's two results
	if (first >= valid) return;
	if (kCompact) {
		uint32_t* out_c = (uint32_t*)out_t_;
		const uint32_t c0 = res.x ? res.x : res.y, c1 = res.z ? res.z : res.w;
		if (out_vec && first + 1 < valid) {
			if (kSys) asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" :: "l"(out_c + first), "r"(c0), "r"(c1) : "memory");
			else      asm volatile("st.global.cs.v2.u32 [%0], {%1,%2};" :: "l"(out_c + first), "r"(c0), "r"(c1) : "memory");
		} else {
			out_c[first] = c0;
			if (first + 1 < valid) out_c[first + 1] = c1;
		}
		return;
	}
	uint2* out_t = (uint2*)out_t_;
	if (out_vec && first + 1 < valid) {
		if (kSys) asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out_t + first), "r"(res.x), "r"(res.y), "r"(res.z), "r"(res.w) : "memory");
		else      asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(out_t + first), "r"(res.x), "r"(res.y), "r"(res.z), "r"(res.w) : "memory");
	} else {
		out_t[first] = make_uint2(res.x, res.y);
		if (first + 1 < valid) out_t[first + 1] = make_uint2(res.z, res.w);
	}
}

// this lane's 16 B of a tile (requests 2*lane, 2*lane + 1); partial tiles read only what exists
template <bool kSys>
__device__ __forceinline__ uint4 warp_tile_load(const uint2* in_t, uint32_t valid, unsigned lane)
{
	uint4 v = make_uint4(0u, 0u, 0u, 0u);
	const uint32_t first = 2 * lane;
	if (first + 1 < valid) {
		if (kSys) asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in_t + first) : "memory");
		else      asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(in_t + first));
	} else if (first < valid) {
		if (kSys) asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(in_t + first) : "memory");
		else      asm volatile("ld.global.cs.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(in_t + first));
	}
	return v;
}

#ifdef GH_DEFINE_KERNELS
template <bool kPairs>
__global__ void __launch_bounds__(256)
search_quad_staged_kernel(const uint2* __restrict__ in, uint2* __restrict__ out,
		const Bucket* __restrict__ table, size_t n, Geom g, Stats* st, unsigned head, int out_bulk)
{
	__shared__ __align__(128) uint2 q_s[2][kTileReq];
	__shared__ __align__(128) uint2 o_s[2][kTileReq];
	__shared__ __align__(8) unsigned long long bar[2];
	const unsigned lane = threadIdx.x & 31u, sub = lane & 3u, grp0 = lane & ~3u, half = sub & 1u;
	const unsigned quad = threadIdx.x >> 2;                      // 0..63: this group's request inside the tile
	uint32_t h1, h2;

	if (head && blockIdx.x == 0 && threadIdx.x < 32) {           // the request in front of the first aligned tile
		const bool live = lane < 4;
		const uint2 q = live ? ld_stream_u2(in) : make_uint2(0u, 0u);
		const uint2 o = quad_probe<kPairs>(table, g, q, live, sub, grp0, half, h1, h2);
		if (lane == 0) {
			st_stream_u2(out, o);
			if (st) { if (h1) atomicAdd(&st->search_hits_b1, 1ULL); if (h2) atomicAdd(&st->search_hits_b2, 1ULL); }
		}
	}
	const uint2* in_a = in + head; uint2* out_a = out + head;
	const size_t n_a = n - head;
	const size_t tiles = (n_a + kTileReq - 1) / kTileReq;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[0])));
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar[1])));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();

	// tile t is "full" when all 64 requests exist: only full tiles go through the copy unit
	auto issue_load = [&](size_t t, int s) {
		const uint32_t b = smem_u32(&bar[s]);
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(kTileReq * 8) : "memory");
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
			:: "r"(smem_u32(&q_s[s][0])), "l"(in_a + t * kTileReq), "r"(kTileReq * 8), "r"(b) : "memory");
	};
	size_t t = blockIdx.x;
	if (threadIdx.x == 0 && t < tiles && (t + 1) * kTileReq <= n_a) issue_load(t, 0);
	uint32_t it = 0, phases = 0;                                 // bit s of phases: parity of stage s's next completion
	bool store_pending = false;                                  // thread 0 only
	for (; t < tiles; t += gridDim.x, it++) {
		const int s = (int)(it & 1u);
		const bool full = (t + 1) * kTileReq <= n_a;
		const size_t tn = t + gridDim.x;
		if (threadIdx.x == 0 && tn < tiles && (tn + 1) * kTileReq <= n_a) issue_load(tn, s ^ 1);
		const size_t i = t * kTileReq + quad;
		const bool live = i < n_a;
		uint2 q = make_uint2(0u, 0u);
		if (full) {
			mbar_wait(smem_u32(&bar[s]), (phases >> s) & 1u);
			phases ^= 1u << s;
			q = q_s[s][quad];
		} else if (live) {
			q = ld_stream_u2(in_a + i);
		}
		const uint2 o = quad_probe<kPairs>(table, g, q, live, sub, grp0, half, h1, h2);
		const bool bulk = full && out_bulk;
		if (live && sub == 0) {
			if (bulk) o_s[s][quad] = o; else st_stream_u2(out_a + i, o);
			if (st) { if (h1) atomicAdd(&st->search_hits_b1, 1ULL); if (h2) atomicAdd(&st->search_hits_b2, 1ULL); }
		}
		// o_s[s] was the source of the bulk store issued two iterations ago; o_s[s^1] of the one issued last
		// iteration: thread 0 makes sure that one has been read out before anyone passes the barrier, because
		// the next iteration writes o_s[s^1] again.
		if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		if (threadIdx.x == 0 && store_pending) { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); store_pending = false; }
		__syncthreads();
		if (threadIdx.x == 0 && bulk) {
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
				:: "l"(out_a + t * kTileReq), "r"(smem_u32(&o_s[s][0])), "r"(kTileReq * 8) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			store_pending = true;
		}
	}
	if (threadIdx.x == 0 && store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Launch-path kernel built on warp_tile_search: every warp walks tiles of 64 requests on its own (grid-stride by warp),
// the next tile's requests already in flight while this one is probed.
template <bool kPairs, bool kCompact = false>
__global__ void __launch_bounds__(256)
search_warp_kernel(const uint2* __restrict__ in, void* __restrict__ out_,
		const Bucket* __restrict__ table, size_t n, Geom g, Stats* st, unsigned head, int out_vec)
{
	constexpr size_t kOutBytes = kCompact ? 4 : 8;               // per request
	char* out = (char*)out_;
	pdl_enter();
	const unsigned lane = threadIdx.x & 31u;
	const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, warps = ((size_t)gridDim.x * blockDim.x) >> 5;
	uint32_t h1 = 0, h2 = 0;
	if (head && warp == 0) {                                     // the request in front of the first 16 B-aligned tile
		const uint4 v = warp_tile_load<false>(in, 1u, lane);
		warp_tile_search<kPairs, false, kCompact>(table, g, in, out, 1u, false, v, lane, h1, h2);
	}
	const uint2* in_a = in + head; char* out_a = out + kOutBytes * head;
	const size_t n_a = n - head;
	const size_t tiles = (n_a + kTileReq - 1) / kTileReq;
	size_t t = warp;
	uint4 v = make_uint4(0u, 0u, 0u, 0u);
	if (t < tiles) v = warp_tile_load<false>(in_a + t * kTileReq, (uint32_t)min((size_t)kTileReq, n_a - t * kTileReq), lane);
	for (; t < tiles; t += warps) {
		const uint32_t valid = (uint32_t)min((size_t)kTileReq, n_a - t * kTileReq);
		const size_t tn = t + warps;
		uint4 vn = make_uint4(0u, 0u, 0u, 0u);
		if (tn < tiles) vn = warp_tile_load<false>(in_a + tn * kTileReq, (uint32_t)min((size_t)kTileReq, n_a - tn * kTileReq), lane);
		warp_tile_search<kPairs, false, kCompact>(table, g, in_a + t * kTileReq, out_a + kOutBytes * t * kTileReq, valid, out_vec != 0, v, lane, h1, h2);
		v = vn;
	}
	if (st) {
		if (h1) atomicAdd(&st->search_hits_b1, (unsigned long long)h1);
		if (h2) atomicAdd(&st->search_hits_b2, (unsigned long long)h2);
	}
}

/* ---- the step in front of the path: key bytes -> (sig, hash) ---- */

// src/mega_recv.c:349-362: sig64 = first 8 key bytes; with -DSIGNATURE the remaining 8-byte words are XOR-folded into it,
// the last, partial word masked to the bytes that belong to the key (:352-359); hash = high 32 bits, sig = low 32 bits.
// Keys of `nkey` >= 8 bytes at keys + i * stride.  (Without SIGNATURE the request IS the first 8 key bytes -- fold = 0.)
// One thread per key, byte loads for the tail so nothing past the key is read (the reference reads the full word and masks).
__global__ void __launch_bounds__(256)
fold_keys_kernel(const unsigned char* __restrict__ keys, size_t stride, uint32_t nkey, int fold, size_t n, uint2* __restrict__ out)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const unsigned char* k = keys + i * stride;
		unsigned long long sig = 0;
#pragma unroll
		for (int b = 0; b < 8; b++) sig |= (unsigned long long)k[b] << (8 * b);
		if (fold) {
			uint32_t j = 8;
			for (; j + 8 <= nkey; j += 8) {
				unsigned long long w = 0;
#pragma unroll
				for (int b = 0; b < 8; b++) w |= (unsigned long long)k[j + b] << (8 * b);
				sig ^= w;
			}
			if (j < nkey) {
				unsigned long long w = 0;
				for (uint32_t b = 0; j + b < nkey; b++) w |= (unsigned long long)k[j + b] << (8 * b);
				sig ^= w;
			}
		}
		out[i] = make_uint2((uint32_t)sig, (uint32_t)(sig >> 32));
	}
}

/* ---- one launch per scheduler cycle, all workers: search -> delete -> insert per worker ---- */

// The reference's cycle walks every worker's batch -- H2D, search launch, D2H, then the delete and the insert launch on
// the worker's stream -- and synchronises ONCE (mega_scheduler.c:393-420, 440-502, 504); it also declares a fused
// gpu_delete_insert it never defines (libgpuhash.h:53-62).  Here the whole cycle of all W workers is ONE launch over a
// table of W batch descriptors.  Work is cut into tiles of 64 requests; the tile index space is phase-major: all search
// tiles (worker 0..W-1), then all delete tiles, then all insert tiles.  Every WARP takes its next tile with an atomic
// ticket, so a tile is only ever held by a warp that is running, and the only things a tile waits for -- the search
// tiles of its own worker (delete), the search and delete tiles of its own worker (insert) -- have smaller indices:
// they are held by running warps or finished.  No assumption about the order in which CTAs are dispatched, about how
// many are resident, or about other kernels sharing the GPU (the first version ordered the phases by blockIdx).
// Per worker the order is the reference's in-stream order search -> delete -> insert; workers are unordered against
// each other, like the reference's streams: a delete tile of worker w waits for w's search tiles only, an insert tile
// for w's search and delete tiles.  In phase-major ticket order those were handed out long before, so the waits are
// short or empty and no SM slot idles through a cycle-wide drain.  A warp publishes finished tiles (fence + one add
// on the worker's counter) when it leaves a claim or a span.  Waits are bounded (timeout -> error word, warp goes on).
//   workspace `ws` (caller-owned, zero before the first launch, left zero by every launch), 32-bit words:
//     [0] tile ticket  [1] CTAs finished  [2] error (sticky, never cleared by the kernel)
//     [32 + 16 w] search tiles of worker w finished   [40 + 16 w] delete tiles of worker w finished
//   (the ticket, hammered by every warp, and each polled counter sit in different 32 B sectors)
struct BatchDesc {                                        // == gpuhash_batch_t (gpuhash_ex.h)
	const void* search_in; void* search_out;              // selem_t[n_search]; loc_t[2 n_search] (compact: loc_t[n_search])
	const void* delete_in; const void* insert_in;         // delem_t[n_delete]; ielem_t[n_insert]
	uint32_t n_search, n_delete, n_insert, reserved;
};
static_assert(sizeof(BatchDesc) == 48, "gpuhash_batch_t");

constexpr int kMaxBatches = 128;                          // descriptors per launch
constexpr int kMaxSegs = 64;                              // legacy insert segments (device-side counts) per launch
constexpr int kMaxSpans = 3 * kMaxBatches + kMaxSegs;

struct MultiArgs {
	const BatchDesc* descs; int W;                        // device-readable descriptor table, or (descs == NULL, W == 1) d0
	BatchDesc d0;
	const uint32_t* const* seg_ptrs; const int* seg_counts; int num_segs;   // extra insert spans of worker 0 (gpu_hash_insert's
	                                                                        // blk_input / blk_elem_num: both live in device memory)
	uint32_t* ws;
	uint32_t upd_tile;                                    // requests per delete / insert tile: 16, 32 or 64 (small cycles take small
	                                                      // tiles so that a lone batch's updates spread over many warps)
	volatile uint32_t* err_host;                          // optional pinned word the host polls after its sync
	unsigned long long timeout_ns;
};

struct SpanTable {                                        // views into dynamic shared memory, built by every CTA
	unsigned char* base; int spans, W;                    // (derived from kernel parameters: nothing here needs a register for long)
	__device__ __forceinline__ const void*& in(int s) const { return ((const void**)base)[s]; }                      // [spans]
	__device__ __forceinline__ void*& out(int w) const { return ((void**)(base + (size_t)spans * 8))[w]; }           // [W]
	__device__ __forceinline__ uint32_t& n(int s) const { return ((uint32_t*)(base + (size_t)spans * 8 + (size_t)W * 8))[s]; }                    // [spans]
	__device__ __forceinline__ uint32_t& first(int s) const { return ((uint32_t*)(base + (size_t)spans * 8 + (size_t)W * 8 + (size_t)spans * 4))[s]; }   // [spans + 1]
};
__host__ __device__ inline size_t span_table_bytes(int W, int segs)
{
	const size_t spans = 3 * (size_t)W + (size_t)segs;
	return spans * 8 + (size_t)W * 8 + spans * 4 + (spans + 1) * 4 + 16;
}

__device__ __forceinline__ uint32_t tiles_of_search(const void* in, uint32_t n)
{
	if (n == 0) return 0;
	const uint32_t head = ((uintptr_t)in & 15u) ? 1u : 0u;   // tiles start at the first 16 B-aligned request
	const uint32_t t = (n - head + kTileReq - 1) / kTileReq;
	return t ? t : 1u;                                    // a lone misaligned request still needs someone to serve it
}

__device__ __forceinline__ void span_table_build(SpanTable& S, unsigned char* smem, const MultiArgs& a)
{
	const int W = a.W, segs = a.num_segs;
	const int spans = 3 * W + segs;
	S.base = smem; S.spans = spans; S.W = W;
	for (int w = threadIdx.x; w < W; w += blockDim.x) {
		const BatchDesc d = a.descs ? a.descs[w] : a.d0;
		S.in(w) = d.search_in; S.out(w) = d.search_out; S.n(w) = d.n_search;
		S.in(W + w) = d.delete_in; S.n(W + w) = d.n_delete;
		S.in(2 * W + w) = d.insert_in; S.n(2 * W + w) = d.n_insert;
	}
	for (int k = threadIdx.x; k < segs; k += blockDim.x) {
		const int c = a.seg_counts[k];
		S.in(3 * W + k) = a.seg_ptrs[k]; S.n(3 * W + k) = c > 0 ? (uint32_t)c : 0u;
	}
	__syncthreads();
	if (threadIdx.x < 32) {                               // exclusive prefix of the tile counts by one warp
		const int per = (spans + 31) / 32, lo = (int)threadIdx.x * per, hi = min(spans, lo + per);
		uint32_t sum = 0;
		const uint32_t ut = a.upd_tile;
		for (int s = lo; s < hi; s++) sum += s < W ? tiles_of_search(S.in(s), S.n(s)) : (S.n(s) + ut - 1) / ut;
		uint32_t inc = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d); if ((int)threadIdx.x >= d) inc += o; }
		uint32_t acc = inc - sum;
		for (int s = lo; s < hi; s++) { S.first(s) = acc; acc += s < W ? tiles_of_search(S.in(s), S.n(s)) : (S.n(s) + ut - 1) / ut; }
		if (threadIdx.x == 31) S.first(spans) = inc;
	}
	__syncthreads();
}

__device__ __forceinline__ int span_of_tile(const SpanTable& S, uint32_t t)          // largest s with first[s] <= t (t < all tiles)
{
	int lo = 0, hi = S.spans;                             // invariant: first[lo] <= t < first[hi]
	while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.first(mid) <= t) lo = mid; else hi = mid; }
	return lo;
}

__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p)
{
	uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// whole warp: wait until *c >= target (lane 0 polls; everyone leaves together)
__device__ __forceinline__ void tiles_wait(const uint32_t* c, uint32_t target, const MultiArgs& a, unsigned lane)
{
	if (target == 0) return;
	if (lane == 0 && ld_acquire_gpu_u32(c) < target) {
		unsigned long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
		unsigned ns = 128u;
		for (;;) {
			__nanosleep(ns);
			if (ld_acquire_gpu_u32(c) >= target) break;
			if (ns < 2048u) ns <<= 1;                      // back off: thousands of warps may be polling this one word
			unsigned long long t1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
			if (t1 - t0 > a.timeout_ns) {                // never seen in practice; a hung box is worse than a flagged batch
				atomicExch(a.ws + 2, 1u);
				if (a.err_host) *a.err_host = 1u;
				break;
			}
		}
	}
	__syncwarp();
}

constexpr uint32_t kWsCounters = 32, kWsPerWorker = 16, kWsDelete = 8;     // word offsets inside the workspace
// Publish `count` finished tiles on counter ws[slot].  What the waiters need ordered before their own table accesses:
//   delete tiles   the CAS writes of the deletes                       -> fence, then add
//   search tiles   nothing but the table READS of the searches, and those have returned their data before the warp gets
//                  here (every lane consumed its rows to form the results; __syncwarp collects the lanes).  The result
//                  stores are for the host / the next kernel, ordered by the kernel boundary.  So: no fence.  A fence here
//                  waits for the warp's result stores to be acknowledged, once per claim: ncu (profiles/r02_xchg_ncu.md)
//                  shows stall_membar as 12 % of a lookup warp's time.  -DGH_PUBLISH_FENCE_ALWAYS restores it for A/B runs.
__device__ __forceinline__ void tiles_publish(uint32_t* ws, uint32_t slot, uint32_t& count, unsigned lane)
{
	if (count == 0) return;                               // warp-uniform
	__syncwarp();
	if (lane == 0) {
#ifdef GH_PUBLISH_FENCE_ALWAYS
		__threadfence();
#else
		if (((slot - kWsCounters) & (kWsPerWorker - 1)) == kWsDelete) __threadfence();
#endif
		atomicAdd(ws + slot, count);
	}
	count = 0;
}
__host__ __device__ inline size_t cycle_workspace_words(int max_batches) { return kWsCounters + (size_t)kWsPerWorker * (size_t)max_batches; }

#ifndef GH_CYCLE_MIN_CTAS
#define GH_CYCLE_MIN_CTAS 2          /* measured, 64 batches of 64 K per launch, mixed / search only, Gops/s: 2 CTAs per SM (128
                                        registers) 19.5 / 19.7, 3 (80 registers, 24-48 B spilled) 17.7 / 17.5, 4 (64) 16.4 / 16.5:
                                        16 warps with four 32 B loads per lane in flight cover the ~47 G probes/s the memory takes */
#endif
constexpr uint32_t kMaxClaim = 4;                         // tiles per ticket at most

template <bool kPairs, bool kCompact>
__global__ void __launch_bounds__(256, GH_CYCLE_MIN_CTAS)
cycle_multi_kernel(Bucket* table, Geom g, Stats* st, MultiArgs a)
{
	extern __shared__ __align__(16) unsigned char span_smem[];
	SpanTable S;                                          // the views live in registers, the arrays in shared memory
	constexpr size_t kOutBytes = kCompact ? 4 : 8;
	pdl_enter();
	span_table_build(S, span_smem, a);
	const unsigned lane = threadIdx.x & 31u;
	const uint32_t total = S.first(S.spans);
	const int W = S.W;
	const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
	uint32_t h1 = 0, h2 = 0;
	uint32_t pend = 0, pend_slot = 0;                     // tiles this warp finished and has not published yet, and their counter

	// Tickets.  A claim takes 1..kMaxClaim consecutive tiles -- many while there is plenty left, one near the end
	// (guided self-scheduling: balance at the tail, few same-address atomics before it).  Claims are issued one chunk
	// ahead of their use and only read (shuffled out of lane 0) when the current chunk runs out; the requests of the
	// NEXT tile are in flight while this one is probed.  Neither latency is on the warp's critical path.
	auto claim_size = [&](uint32_t seen) -> uint32_t {
		const uint32_t rem = total > seen ? total - seen : 0u;
		return min(kMaxClaim, max(1u, rem / (6u * nwarps)));
	};
	auto claim_issue = [&](uint32_t k) -> uint32_t { return lane == 0 ? atomicAdd(a.ws + 0, k) : 0u; };
	// requests of a search tile (this lane's 16 B); nothing for other tiles
	auto prefetch = [&](uint32_t t, int s) -> uint4 {
		if (t >= total || s >= W) return make_uint4(0u, 0u, 0u, 0u);
		const uint2* in = (const uint2*)S.in(s);
		const uint32_t head = ((uintptr_t)in & 15u) ? 1u : 0u;
		const uint32_t n_a = S.n(s) - head, r0 = (t - S.first(s)) * kTileReq;
		return warp_tile_load<false>(in + head + r0, min((uint32_t)kTileReq, n_a - min(n_a, r0)), lane);
	};

	uint32_t k_cur = claim_size(0), k_nxt = k_cur;
	const uint32_t raw_a = claim_issue(k_cur), raw_b = claim_issue(k_nxt);           // two atomics in flight together
	uint32_t c0 = __shfl_sync(0xffffffffu, raw_a, 0), c1 = c0 + k_cur;
	uint32_t n0 = __shfl_sync(0xffffffffu, raw_b, 0), n1 = n0 + k_nxt;
	uint32_t seen = n1, raw_nn = 0, k_nn = 1;
	uint32_t t = c0;
	int s_cur = t < total ? span_of_tile(S, t) : 0;
	uint4 v = prefetch(t, s_cur);
	while (t < total) {
		if (t == c0) { k_nn = claim_size(seen); raw_nn = claim_issue(k_nn); }        // the chunk after next
		const bool last_of_chunk = t + 1 >= c1;
		const uint32_t tn = last_of_chunk ? n0 : t + 1;
		int s_nxt = s_cur;
		if (tn < total) {
			if (last_of_chunk) s_nxt = span_of_tile(S, tn);
			else while (tn >= S.first(s_nxt + 1)) s_nxt++;
		}
		const uint4 vn = prefetch(tn, s_nxt);             // in flight while this tile is probed
		const int s = s_cur;
		const uint32_t lt = t - S.first(s);               // tile inside its span
		if (s < W) {                                      // ---- search
			const uint2* in = (const uint2*)S.in(s); char* out = (char*)S.out(s);
			const uint32_t n = S.n(s);
			const uint32_t head = ((uintptr_t)in & 15u) ? 1u : 0u;
			const bool out_vec = (((uintptr_t)out + kOutBytes * head) & (kCompact ? 7u : 15u)) == 0;
			if (head && lt == 0) {                        // the request in front of the first aligned tile
				const uint4 v0 = warp_tile_load<false>(in, 1u, lane);
				warp_tile_search<kPairs, false, kCompact>(table, g, in, out, 1u, false, v0, lane, h1, h2);
			}
			const uint32_t n_a = n - head, r0 = lt * kTileReq;
			const uint32_t valid = min((uint32_t)kTileReq, n_a - min(n_a, r0));
			if (valid)
				warp_tile_search<kPairs, false, kCompact>(table, g, in + head + r0, out + kOutBytes * (head + (size_t)r0), valid, out_vec, v, lane, h1, h2);
			const uint32_t slot = kWsCounters + kWsPerWorker * (uint32_t)s;
			if (slot != pend_slot) { tiles_publish(a.ws, pend_slot, pend, lane); pend_slot = slot; }
			pend++;
		} else {                                          // ---- delete / insert: after the searches (and deletes) of ITS worker
			const bool is_delete = s < 2 * W;
			const int w = is_delete ? s - W : (s < 3 * W ? s - 2 * W : 0);       // segments belong to worker 0
			tiles_publish(a.ws, pend_slot, pend, lane);   // this warp's own finished tiles first: nobody waits on a waiter
			tiles_wait(a.ws + kWsCounters + kWsPerWorker * w, S.first(w + 1) - S.first(w), a, lane);
			if (!is_delete) tiles_wait(a.ws + kWsCounters + kWsPerWorker * w + kWsDelete, S.first(W + w + 1) - S.first(W + w), a, lane);
			const uint32_t* in = (const uint32_t*)S.in(s);
			const uint32_t n = S.n(s), r0 = lt * a.upd_tile;
			if (kPairs) {                                 // two lanes per request: 16 requests per round
#pragma unroll 1
				for (uint32_t r = 0; r < a.upd_tile; r += 16) {
					if (r0 + r >= n) break;               // warp-uniform
					const uint32_t i = r0 + r + (lane >> 1);
					const bool have = i < n;
					uint32_t x = 0, y = 0, z = 0;
					if (have) { x = ld_stream_u32(in + 3 * (size_t)i); y = ld_stream_u32(in + 3 * (size_t)i + 1); z = ld_stream_u32(in + 3 * (size_t)i + 2); }
					if (is_delete) {
						const int zc = delete_pair(table, g, have, x, y, z, lane);
						if (st && zc && (lane & 1u) == 0) { atomicAdd(&st->del_zeroed, (unsigned long long)zc); atomicAdd(&st->del_requests_hit, 1ULL); }
					} else {
						insert_pair(table, g, have, x, y, z, st, lane);
					}
				}
			} else {
#pragma unroll 1
				for (uint32_t r = 0; r < a.upd_tile; r += 32) {
					const uint32_t i = r0 + r + lane;
					if (i < n && r + lane < a.upd_tile) {
						const uint32_t x = ld_stream_u32(in + 3 * (size_t)i), y = ld_stream_u32(in + 3 * (size_t)i + 1), z = ld_stream_u32(in + 3 * (size_t)i + 2);
						if (is_delete) {
							const int zc = delete_one<false>(table, g, x, y, z);
							if (st && zc) { atomicAdd(&st->del_zeroed, (unsigned long long)zc); atomicAdd(&st->del_requests_hit, 1ULL); }
						} else {
							insert_one<false>(table, g, x, y, z, st);
						}
					}
				}
			}
			if (is_delete) { pend_slot = kWsCounters + kWsPerWorker * (uint32_t)w + kWsDelete; pend = 1; }
		}
		if (last_of_chunk) {
			tiles_publish(a.ws, pend_slot, pend, lane);   // bounded delay: others may be waiting for exactly these tiles
			c0 = n0; c1 = n1;
			n0 = __shfl_sync(0xffffffffu, raw_nn, 0); n1 = n0 + k_nn;
			seen = max(seen, n1);
		}
		t = tn; s_cur = s_nxt; v = vn;
	}
	tiles_publish(a.ws, pend_slot, pend, lane);
	if (st) {                                             // every quad's first lane counted its own hits
		if (h1) atomicAdd(&st->search_hits_b1, (unsigned long long)h1);
		if (h2) atomicAdd(&st->search_hits_b2, (unsigned long long)h2);
	}
	// ---- this CTA is done; the last one leaves the workspace zero for its next user (the error word stays)
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		if (atomicAdd(a.ws + 1, 1u) == gridDim.x - 1) {
			a.ws[0] = 0; a.ws[1] = 0;
			for (int w = 0; w < W; w++) { a.ws[kWsCounters + kWsPerWorker * w] = 0; a.ws[kWsCounters + kWsPerWorker * w + kWsDelete] = 0; }
			__threadfence();
		}
	}
}

/* ---- alternative search shape, kept for comparison (tools/sweep.py) ---- */

__device__ __forceinline__ uint4 ld_half_row(const uint32_t* p)
{
	uint4 v;
	asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		: "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// Split layout only: four lanes per request -- lanes 0,1 take the two 16 B halves of bucket 1's
// signature row, lanes 2,3 those of bucket 2 -- 128-bit loads, ballot inside the 4-lane group,
// the hit lane fetches the location, lane 0 stores the pair.  Same results as search_kernel.
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
			uint4 v = ld_half_row(table[b].w + 4 * (sub & 1u));
			m = (v.x == q.x ? 1u : 0u) | (v.y == q.x ? 2u : 0u) | (v.z == q.x ? 4u : 0u) | (v.w == q.x ? 8u : 0u);
			if (m) loc = ld_u32_ro(&table[b].w[8 + 4 * (sub & 1u) + (__ffs(m) - 1)]);
		}
		const unsigned hits = (__ballot_sync(0xffffffffu, m != 0) >> grp0) & 0xfu;
		const uint32_t l0 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 1u) ? 0 : 1));   // lower half wins
		const uint32_t l1 = __shfl_sync(0xffffffffu, loc, grp0 + ((hits & 4u) ? 2 : 3));
		if (live && sub == 0)
			st_stream_u2(out + i, make_uint2((hits & 3u) ? l0 : 0u, (hits & 12u) ? l1 : 0u));
	}
}
#endif  /* GH_DEFINE_KERNELS */

}  // namespace gh
