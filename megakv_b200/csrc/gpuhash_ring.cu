/*
 * gpuhash_ring.cu -- the scheduler cycle without launches: pinned-host descriptor rings feeding ONE persistent kernel
 * (BASELINE.json north_star (c)).
 *
 * Reference: every 200 us the scheduler thread issues, per worker, 2-3 copies and 3 kernel launches and then blocks in
 * cudaDeviceSynchronize (src/mega_scheduler.c:392-504).  Here the host side of a cycle is one 64-byte descriptor
 * written into a ring in pinned memory -- pointers to the worker's own pinned batch buffers (the ones
 * src/mega_recv.c:154-156,176 allocates), the three counts and, last, a sequence number (the doorbell).  No CUDA call.
 *
 *   ring r (one per worker, or several workers per ring)        persistent kernel, R groups of CTAs
 *   host: desc[slot] = {ptrs, counts}; desc[slot].seq = b  -->  group r, leader thread polls desc[b % slots].seq over PCIe,
 *                                                               copies the descriptor into device memory, raises `go`
 *                                                               all warps of the group: searches (64-request tiles, the
 *                                                               requests pulled from and the results pushed to the pinned
 *                                                               buffers as 512 B system-scope vector accesses), group
 *                                                               barrier, deletes, group barrier, inserts
 *   host: spins on done[slot] == b                         <--  last CTA: done[slot] = b (release, system scope)
 *
 * Batches of one ring are processed strictly in order and in the reference's in-batch order search -> delete -> insert;
 * rings are unordered against each other -- exactly the per-stream ordering of the reference (mega_scheduler.c:392-502).
 * Request and result bytes never touch a staging buffer: H2D, lookup and D2H of different tiles and different rings
 * overlap inside the kernel.
 *
 * A persistent kernel has no launch boundary at which the host's writes become visible, so visibility is built from
 * the memory model: doorbell and completion mark are system-scope release/acquire, `go` is device-scope
 * release/acquire, and the request/result bytes in between are ordinary (coalescing) accesses ordered by that chain.
 * tests/test_gpu_ring.py refills the SAME pinned buffers batch after batch to catch a stale read.
 *
 * Safety: the kernel parks itself (every group leaves at its next batch boundary) when the host asks for it or when no
 * doorbell rang on any ring for idle_ms; the next submit (or a wait that finds the kernel parked) relaunches it and it
 * resumes at the first batch whose `done` is missing.  One ring object per device.  It never waits on anything another
 * resident CTA cannot provide: the grid is capped to what is co-resident (occupancy API), so the group barriers cannot
 * deadlock.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <time.h>
#include <cuda_runtime.h>

#include "gpuhash_ex.h"
#include "gpuhash_kernels.cuh"

#define RING_MAX 32

struct __align__(64) RingDesc {                 /* pinned host memory, one per slot */
	const void *search_in; void *search_out;    /* selem_t[n_search], loc_t[2 n_search] (pinned, device-visible) */
	const void *delete_in; const void *insert_in;
	uint32_t n_search, n_delete, n_insert;
	volatile uint32_t seq;                      /* doorbell: batch number b (1-based), written last */
	volatile uint32_t done;                     /* written by the GPU: b when batch b is complete */
	uint32_t pad[3];
};
static_assert(sizeof(RingDesc) == 64, "descriptor is one 64 B line");

struct GroupCtl {                               /* device memory, one per ring */
	uint32_t go;                                /* batch number the group may work on (0xffffffff: exit) */
	uint32_t n_search, n_delete, n_insert;
	const void *search_in; void *search_out; const void *delete_in; const void *insert_in;
	uint32_t arrive[3];                         /* monotonic: phase barriers and completion count */
	uint32_t pad;
};

struct RingGlobal {                             /* device memory, one per kernel */
	unsigned long long last_activity;           /* globaltimer of the last doorbell any leader saw */
	uint32_t parking;                           /* a leader decided to park: every group leaves at its next batch boundary */
	uint32_t pad;
};

struct RingParams {
	RingDesc *desc;                             /* [rings][slots] (device-visible address of the pinned array) */
	GroupCtl *ctl;                              /* [rings] */
	RingGlobal *glob;
	volatile uint32_t *host_flags;              /* pinned: [0] host sets 1 to park the kernel, [1] kernel sets 1 when it parks */
	uint32_t *first_batch;                      /* [rings] device: batch number each group starts with (relaunch) */
	unsigned long long *trace;                  /* [rings][8] device, globaltimer stamps of the latest batch (diagnosis) */
	int rings, slots, ctas_per_ring;
	unsigned long long idle_ns;
};

namespace {

__device__ __forceinline__ uint32_t ld_sys_u32(const volatile uint32_t *p)
{
	uint32_t v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint4 ld_sys_u4(const void *p)
{
	uint4 v; asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint32_t ld_gpu_acquire(const uint32_t *p)
{
	uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ uint64_t now_ns() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

/* all CTAs of a group have arrived `target` times at counter c (monotonic, so nothing is ever reset) */
__device__ __forceinline__ void group_barrier(uint32_t *c, uint32_t target)
{
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		atomicAdd(c, 1u);
		while (ld_gpu_acquire(c) < target) __nanosleep(100);
	}
	__syncthreads();
}

/* kSysData: request/result bytes with system-scope accesses instead of weak ones.  Weak accesses are what the memory
 * model needs here -- host data -> host release(seq) -> leader: relaxed.sys load + fence.acq_rel.sys -> release.gpu(go)
 * -> worker: acquire.gpu(go) -> weak load of the data is a causality chain, and results -> __threadfence_system ->
 * release.sys(done) the same in reverse -- and they coalesce into 512 B transactions, which system-scope vector
 * accesses do not (measured: lone-batch round trip 70 us with them).  Kept as a compile-time switch for diagnosis. */
#ifndef GPUHASH_RING_SYS_DATA
#define GPUHASH_RING_SYS_DATA 0
#endif
constexpr bool kSysData = GPUHASH_RING_SYS_DATA != 0;

template <bool kPairs>
__global__ void __launch_bounds__(256)
ring_kernel(gh::Bucket *table, gh::Geom g, RingParams P)
{
	__shared__ uint32_t s_go, s_ns, s_nd, s_ni;
	__shared__ const void *s_sin, *s_din, *s_iin; __shared__ void *s_sout;
	const int r = blockIdx.x / P.ctas_per_ring, cta = blockIdx.x % P.ctas_per_ring, cpr = P.ctas_per_ring;
	GroupCtl *ctl = P.ctl + r;
	const unsigned lane = threadIdx.x & 31u;
	uint32_t b = P.first_batch[r];                                /* next batch number of this ring */
	const uint32_t b0 = b;

	for (;; b++) {
		/* ---- the leader watches the doorbell over PCIe; everybody else watches `go` in device memory ---- */
		if (cta == 0 && threadIdx.x == 0) {
			/* batch b-1 must be complete in every CTA of the group before its descriptor is replaced (and before a
			 * search of batch b may run: in-ring order is strict) */
			if (b == b0) atomicMax(&P.glob->last_activity, (unsigned long long)now_ns());      /* the launch counts as activity */
			while (ld_gpu_acquire(&ctl->arrive[2]) < (uint32_t)cpr * (b - b0)) __nanosleep(100);
			RingDesc *d = P.desc + (size_t)r * P.slots + (b - 1) % P.slots;
			unsigned long long *tr = P.trace + 8 * r;
			tr[5] = now_ns();
			uint32_t go = b;
			uint4 p0, p1, c;
			/* one 4-byte read over the link per poll; the rare exits (host stop request, idle timeout) every 16th */
			for (unsigned spin = 0;; spin++) {
				if (ld_gpu_acquire(&P.glob->parking) != 0u) { go = 0xffffffffu; break; }
				if (ld_sys_u32(&d->seq) == b) {
					asm volatile("fence.acq_rel.sys;" ::: "memory");      /* the descriptor was written before the doorbell */
					p0 = ld_sys_u4(d); p1 = ld_sys_u4((const char *)d + 16); c = ld_sys_u4((const char *)d + 32);
					*(volatile unsigned long long *)&P.glob->last_activity = now_ns();
					break;
				}
				if ((spin & 15u) == 15u) {
					if (ld_sys_u32(P.host_flags) != 0u) { go = 0xffffffffu; break; }
					if (now_ns() - *(volatile unsigned long long *)&P.glob->last_activity > P.idle_ns) { go = 0xffffffffu; break; }
				}
				__nanosleep(100);
			}
			if (go == b) {
				tr[0] = now_ns(); tr[2] = 0; tr[4] = 0;
				ctl->search_in = (const void *)(((uint64_t)p0.y << 32) | p0.x); ctl->search_out = (void *)(((uint64_t)p0.w << 32) | p0.z);
				ctl->delete_in = (const void *)(((uint64_t)p1.y << 32) | p1.x); ctl->insert_in = (const void *)(((uint64_t)p1.w << 32) | p1.z);
				ctl->n_search = c.x; ctl->n_delete = c.y; ctl->n_insert = c.z;
				tr[1] = now_ns();
			} else {
				atomicExch(&P.glob->parking, 1u);                 /* collective: nobody keeps serving while others have left */
				asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(P.host_flags + 1), "r"(1u) : "memory");
				__threadfence_system();
			}
			asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(&ctl->go), "r"(go) : "memory");
		}
		if (threadIdx.x == 0) {
			uint32_t go;
			unsigned ns = 100;
			while ((go = ld_gpu_acquire(&ctl->go)) != 0xffffffffu && (int32_t)(go - b) < 0) { __nanosleep(ns); if (ns < 1000) ns += 100; }
			s_go = go;
			if (go != 0xffffffffu) atomicMax(P.trace + 8 * r + 4, (unsigned long long)now_ns());
			if (go != 0xffffffffu) {
				s_ns = ctl->n_search; s_nd = ctl->n_delete; s_ni = ctl->n_insert;
				s_sin = ctl->search_in; s_sout = ctl->search_out; s_din = ctl->delete_in; s_iin = ctl->insert_in;
			}
		}
		__syncthreads();
		if (s_go == 0xffffffffu) break;
		const uint32_t nb = b - b0 + 1;                           /* batches this launch has seen, for the monotonic counters */
		const uint32_t n_search = s_ns, n_delete = s_nd, n_insert = s_ni;

		/* ---- searches: every warp of the group walks 64-request tiles on its own (gh::warp_tile_search): one 512 B
		 *      system-scope read from the pinned request buffer (the next tile's already in flight), four table loads
		 *      per lane in flight, one 512 B write into the pinned result buffer.  No shared memory, no CTA barrier. ---- */
		if (n_search) {
			const uint2 *in = (const uint2 *)s_sin; uint2 *out = (uint2 *)s_sout;
			const uint32_t head = ((uintptr_t)in & 15u) ? 1u : 0u;        /* tiles start at the first 16 B-aligned request */
			const bool out_vec = (((uintptr_t)out + 8u * head) & 15u) == 0;
			const uint32_t warp = (uint32_t)cta * 8u + (threadIdx.x >> 5), warps = (uint32_t)cpr * 8u;
			uint32_t h1 = 0, h2 = 0;
			if (head && warp == 0) {
				const uint4 v = gh::warp_tile_load<kSysData>(in, 1u, lane);
				gh::warp_tile_search<kPairs, kSysData>(table, g, in, out, 1u, false, v, lane, h1, h2);
			}
			const uint2 *in_a = in + head; uint2 *out_a = out + head;
			const uint32_t n_a = n_search - head;
			const uint32_t tiles = (n_a + gh::kTileReq - 1) / gh::kTileReq;
			uint32_t t = warp;
			uint4 v = make_uint4(0u, 0u, 0u, 0u);
			if (t < tiles) v = gh::warp_tile_load<kSysData>(in_a + (size_t)t * gh::kTileReq, min((uint32_t)gh::kTileReq, n_a - t * gh::kTileReq), lane);
			for (; t < tiles; t += warps) {
				const uint32_t valid = min((uint32_t)gh::kTileReq, n_a - t * gh::kTileReq);
				const uint32_t tn = t + warps;
				uint4 vn = make_uint4(0u, 0u, 0u, 0u);
				if (tn < tiles) vn = gh::warp_tile_load<kSysData>(in_a + (size_t)tn * gh::kTileReq, min((uint32_t)gh::kTileReq, n_a - tn * gh::kTileReq), lane);
				gh::warp_tile_search<kPairs, kSysData>(table, g, in_a + (size_t)t * gh::kTileReq, out_a + (size_t)t * gh::kTileReq, valid, out_vec, v, lane, h1, h2);
				v = vn;
			}
		}
		/* ---- deletes after every search of the batch, inserts after every delete (gpu_hash.cu order per stream) ---- */
		if (n_delete) {
			group_barrier(&ctl->arrive[0], (uint32_t)cpr * nb);
			const uint32_t *in = (const uint32_t *)s_din;
			if (kPairs) {                                                 /* two lanes per request: the bucket is one L2 request */
				const uint32_t n_up = (n_delete + 15u) & ~15u;
				for (uint32_t i = (cta * blockDim.x + threadIdx.x) >> 1; i < n_up; i += (cpr * blockDim.x) >> 1) {
					const bool have = i < n_delete;
					uint32_t a = 0, bb = 0, c = 0;
					if (have) { a = gh::ld_stream_u32(in + 3 * i); bb = gh::ld_stream_u32(in + 3 * i + 1); c = gh::ld_stream_u32(in + 3 * i + 2); }
					gh::delete_pair(table, g, have, a, bb, c, lane);
				}
			} else
			for (uint32_t i = cta * blockDim.x + threadIdx.x; i < n_delete; i += cpr * blockDim.x)
				gh::delete_one<kPairs>(table, g, gh::ld_stream_u32(in + 3 * i), gh::ld_stream_u32(in + 3 * i + 1), gh::ld_stream_u32(in + 3 * i + 2));
		} else if (threadIdx.x == 0) atomicAdd(&ctl->arrive[0], 1u);      /* counters stay in step with nb */
		if (n_insert) {
			group_barrier(&ctl->arrive[1], (uint32_t)cpr * nb);
			const uint32_t *in = (const uint32_t *)s_iin;
			if (kPairs) {
				const uint32_t n_up = (n_insert + 15u) & ~15u;
				for (uint32_t i = (cta * blockDim.x + threadIdx.x) >> 1; i < n_up; i += (cpr * blockDim.x) >> 1) {
					const bool have = i < n_insert;
					uint32_t a = 0, bb = 0, c = 0;
					if (have) { a = gh::ld_stream_u32(in + 3 * i); bb = gh::ld_stream_u32(in + 3 * i + 1); c = gh::ld_stream_u32(in + 3 * i + 2); }
					gh::insert_pair(table, g, have, a, bb, c, nullptr, lane);
				}
			} else
			for (uint32_t i = cta * blockDim.x + threadIdx.x; i < n_insert; i += cpr * blockDim.x)
				gh::insert_one<kPairs>(table, g, gh::ld_stream_u32(in + 3 * i), gh::ld_stream_u32(in + 3 * i + 1), gh::ld_stream_u32(in + 3 * i + 2), nullptr);
		} else if (threadIdx.x == 0) atomicAdd(&ctl->arrive[1], 1u);
		/* ---- completion: the last CTA of the group tells the host.  Every CTA orders its result stores before its
		 *      arrival at device scope; the last one to arrive orders all of that before the completion mark at system
		 *      scope (fences are cumulative), so only one CTA per batch pays for a system-scope fence. ---- */
		__syncthreads();
		if (threadIdx.x == 0) {
			atomicMax(P.trace + 8 * r + 2, (unsigned long long)now_ns());
			__threadfence();
			if (atomicAdd(&ctl->arrive[2], 1u) == (uint32_t)cpr * nb - 1u) {
				RingDesc *d = P.desc + (size_t)r * P.slots + (b - 1) % P.slots;
				__threadfence_system();
				asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(&d->done), "r"(b) : "memory");
				P.trace[8 * r + 3] = now_ns();
			}
		}
	}
}

}  // namespace

struct gpuhash_ring_s {
	gpuhash_geom_t geom;
	void *table;
	int rings, slots, ctas_per_ring;
	RingDesc *desc_h; RingDesc *desc_d;           /* same pinned array, host and device view */
	GroupCtl *ctl_d;
	uint32_t *flags_h, *flags_d;                  /* pinned: [0] stop request (host), [1] parked (kernel) */
	RingGlobal *glob_d;
	unsigned long long *trace_d;
	uint32_t *first_d;
	int host_ptr_ok;                              /* device address of pinned memory == host address (UVA) */
	uint32_t next[RING_MAX];                      /* next batch number to submit, per ring (1-based) */
	cudaStream_t stream;
	int running;
	unsigned idle_ms;
};

static double wall_ms(void)
{
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

static int ring_launch(gpuhash_ring_t *q)
{
	/* every group resumes at the batch it has not completed yet: done[] of the previous batch is the truth */
	uint32_t first[RING_MAX];
	for (int r = 0; r < q->rings; r++) {
		uint32_t b = 1;
		for (int s = 0; s < q->slots; s++) { uint32_t d = q->desc_h[(size_t)r * q->slots + s].done; if (d + 1 > b) b = d + 1; }
		first[r] = b;
	}
	cudaError_t e = cudaMemcpyAsync(q->first_d, first, sizeof(uint32_t) * q->rings, cudaMemcpyHostToDevice, q->stream);
	if (e != cudaSuccess) return (int)e;
	if ((e = cudaMemsetAsync(q->ctl_d, 0, sizeof(GroupCtl) * q->rings, q->stream)) != cudaSuccess) return (int)e;
	if ((e = cudaMemsetAsync(q->glob_d, 0, sizeof(RingGlobal), q->stream)) != cudaSuccess) return (int)e;
	q->flags_h[0] = 0; q->flags_h[1] = 0;
	RingParams P;
	P.desc = q->desc_d; P.ctl = q->ctl_d; P.glob = q->glob_d; P.host_flags = q->flags_d; P.first_batch = q->first_d; P.trace = q->trace_d;
	P.rings = q->rings; P.slots = q->slots; P.ctas_per_ring = q->ctas_per_ring;
	P.idle_ns = (unsigned long long)q->idle_ms * 1000000ULL;
	gh::Geom gg; gg.hash_mask = q->geom.hash_mask; gg.block_mask = q->geom.block_mask; gg.algo = q->geom.algo;
	gg.max_cuckoo = q->geom.max_cuckoo; gg.layout = q->geom.layout;
	const unsigned grid = (unsigned)(q->rings * q->ctas_per_ring);
	if (gg.layout == gh::kLayoutPairs) ring_kernel<true><<<grid, 256, 0, q->stream>>>((gh::Bucket *)q->table, gg, P);
	else                               ring_kernel<false><<<grid, 256, 0, q->stream>>>((gh::Bucket *)q->table, gg, P);
	e = cudaGetLastError();
	if (e == cudaSuccess) q->running = 1;
	return (int)e;
}

extern "C" gpuhash_ring_t *gpuhash_ring_create(const gpuhash_geom_t *g, void *table_d, int rings, int slots,
		int ctas_per_sm, unsigned idle_ms)
{
	if (!g || !table_d || rings < 1 || rings > RING_MAX || slots < 1 || slots > 64 || g->layout > GPUHASH_LAYOUT_REFERENCE) return NULL;
	gpuhash_ring_t *q = (gpuhash_ring_t *)calloc(1, sizeof *q);
	if (!q) return NULL;
	q->geom = *g; q->table = table_d; q->rings = rings; q->slots = slots; q->idle_ms = idle_ms ? idle_ms : 2000;
	for (int r = 0; r < rings; r++) q->next[r] = 1;
	/* grid = what is co-resident, so that the group barriers cannot wait for a CTA that is not scheduled */
	int dev = 0, sms = 148, occ = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaError_t e = g->layout == GPUHASH_LAYOUT_PAIRS
		? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel<true>, 256, 0)
		: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ring_kernel<false>, 256, 0);
	if (e != cudaSuccess || occ < 1) { free(q); return NULL; }
	if (ctas_per_sm < 1) ctas_per_sm = 4;                 /* leaves half of every SM to kernels of the legacy entry points */
	if (ctas_per_sm > occ) ctas_per_sm = occ;
	q->ctas_per_ring = sms * ctas_per_sm / rings;
	if (q->ctas_per_ring < 1) { free(q); return NULL; }
	int ok = cudaHostAlloc((void **)&q->desc_h, sizeof(RingDesc) * rings * slots, cudaHostAllocMapped) == cudaSuccess
	      && cudaHostGetDevicePointer((void **)&q->desc_d, q->desc_h, 0) == cudaSuccess
	      && cudaHostAlloc((void **)&q->flags_h, 64, cudaHostAllocMapped) == cudaSuccess
	      && cudaHostGetDevicePointer((void **)&q->flags_d, q->flags_h, 0) == cudaSuccess
	      && cudaMalloc((void **)&q->ctl_d, sizeof(GroupCtl) * rings) == cudaSuccess
	      && cudaMalloc((void **)&q->glob_d, sizeof(RingGlobal)) == cudaSuccess
	      && cudaMalloc((void **)&q->trace_d, sizeof(unsigned long long) * 8 * RING_MAX) == cudaSuccess
	      && cudaMemset(q->trace_d, 0, sizeof(unsigned long long) * 8 * RING_MAX) == cudaSuccess
	      && cudaMalloc((void **)&q->first_d, sizeof(uint32_t) * rings) == cudaSuccess
	      && cudaStreamCreateWithFlags(&q->stream, cudaStreamNonBlocking) == cudaSuccess;
	if (!ok) { fprintf(stderr, "gpuhash_ring_create: %s\n", cudaGetErrorString(cudaGetLastError())); gpuhash_ring_destroy(q); return NULL; }
	memset(q->desc_h, 0, sizeof(RingDesc) * rings * slots);
	memset(q->flags_h, 0, 64);
	{ int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrCanUseHostPointerForRegisteredMem, dev); q->host_ptr_ok = v; }
	if (ring_launch(q) != 0) { gpuhash_ring_destroy(q); return NULL; }
	return q;
}

/* park the kernel: it finishes the batches whose doorbell it has seen, then exits */
extern "C" int gpuhash_ring_park(gpuhash_ring_t *q)
{
	if (!q) return -1;
	if (!q->running) return 0;
	__atomic_store_n(&q->flags_h[0], 1u, __ATOMIC_RELEASE);
	cudaError_t e = cudaStreamSynchronize(q->stream);
	q->running = 0;
	return (int)e;
}

extern "C" void gpuhash_ring_destroy(gpuhash_ring_t *q)
{
	if (!q) return;
	if (q->stream) { gpuhash_ring_park(q); cudaStreamDestroy(q->stream); }
	if (q->desc_h) cudaFreeHost(q->desc_h);
	if (q->flags_h) cudaFreeHost(q->flags_h);
	cudaFree(q->ctl_d); cudaFree(q->first_d); cudaFree(q->glob_d); cudaFree(q->trace_d);
	free(q);
}

/* One scheduler cycle of one worker: same arguments as gpuhash_index_submit, buffers PINNED (cudaHostAlloc /
 * cudaHostRegister).  Returns the batch number (> 0) to wait for, or a negative error.  Blocks only while the ring is
 * full (the slot's previous batch not completed). */
/* The kernel parks itself (idle timeout, gpuhash_ring_park) at a batch boundary, possibly with batches still pending.
 * Whoever finds the parked flag while waiting for one of them lets the kernel finish its exit and launches it again: it
 * resumes at the first batch without a completion mark.  0, or a CUDA / launch error. */
static int ring_relaunch_if_parked(gpuhash_ring_t *q)
{
	if (q->running && !__atomic_load_n(&q->flags_h[1], __ATOMIC_ACQUIRE)) return 0;
	cudaError_t e = cudaStreamSynchronize(q->stream);
	if (e != cudaSuccess) return (int)e;
	q->running = 0;
	return ring_launch(q);
}

extern "C" long long gpuhash_ring_submit(gpuhash_ring_t *q, int ring,
		const void *search_in_h, size_t n_search, void *search_out_h,
		const void *delete_in_h, size_t n_delete, const void *insert_in_h, size_t n_insert)
{
	if (!q || ring < 0 || ring >= q->rings || n_search > 0xffffffffu || n_delete > 0xffffffffu || n_insert > 0xffffffffu) return -1;
	if ((n_search && (!search_in_h || !search_out_h)) || (n_delete && !delete_in_h) || (n_insert && !insert_in_h)) return -1;
	if (((uintptr_t)search_in_h | (uintptr_t)search_out_h) & 7u) return -1;
	int rc = ring_relaunch_if_parked(q);                                 /* parked (idle timeout or gpuhash_ring_park): start it again */
	if (rc) return -(long long)(rc > 0 ? rc : -rc);
	const uint32_t b = q->next[ring];
	RingDesc *d = q->desc_h + (size_t)ring * q->slots + (b - 1) % q->slots;
	if (b > (uint32_t)q->slots) {                                        /* slot still in flight? */
		const double t0 = wall_ms();
		while (__atomic_load_n(&d->done, __ATOMIC_ACQUIRE) != b - (uint32_t)q->slots) {
			if (wall_ms() - t0 > 10000.0) return -2;
			/* the kernel may have parked with this very slot pending: without a relaunch nothing would ever complete it */
			if ((rc = ring_relaunch_if_parked(q)) != 0) return -(long long)(rc > 0 ? rc : -rc);
		}
	}
	d->search_in = search_in_h; d->search_out = search_out_h; d->delete_in = delete_in_h; d->insert_in = insert_in_h;
	if (!q->host_ptr_ok) {                                               /* no unified addressing of pinned memory: translate */
		void *p;
		if (n_search) {
			if (cudaHostGetDevicePointer(&p, (void *)search_in_h, 0) != cudaSuccess) { cudaGetLastError(); return -3; }
			d->search_in = p;
			if (cudaHostGetDevicePointer(&p, search_out_h, 0) != cudaSuccess) { cudaGetLastError(); return -3; }
			d->search_out = p;
		}
		if (n_delete) { if (cudaHostGetDevicePointer(&p, (void *)delete_in_h, 0) != cudaSuccess) { cudaGetLastError(); return -3; } d->delete_in = p; }
		if (n_insert) { if (cudaHostGetDevicePointer(&p, (void *)insert_in_h, 0) != cudaSuccess) { cudaGetLastError(); return -3; } d->insert_in = p; }
	}
	d->n_search = (uint32_t)n_search; d->n_delete = (uint32_t)n_delete; d->n_insert = (uint32_t)n_insert;
	__atomic_store_n(&d->seq, b, __ATOMIC_RELEASE);                      /* the doorbell */
	q->next[ring] = b + 1;
	return (long long)b;
}

/* spin until batch `ticket` of `ring` is complete (its results are in the caller's search_out buffer) */
extern "C" int gpuhash_ring_wait(gpuhash_ring_t *q, int ring, long long ticket, unsigned timeout_ms)
{
	if (!q || ring < 0 || ring >= q->rings || ticket < 1 || (uint32_t)ticket >= q->next[ring]) return -1;
	RingDesc *d = q->desc_h + (size_t)ring * q->slots + ((uint32_t)ticket - 1) % q->slots;
	const double t0 = wall_ms();
	for (;;) {
		const uint32_t done = __atomic_load_n(&d->done, __ATOMIC_ACQUIRE);
		if ((int32_t)(done - (uint32_t)ticket) >= 0) return 0;
		if (timeout_ms && wall_ms() - t0 > (double)timeout_ms) return -2;
		int rc = ring_relaunch_if_parked(q);                                     /* the kernel parked with this batch pending */
		if (rc) return rc;
	}
}

/* everything submitted so far, on every ring */
extern "C" int gpuhash_ring_drain(gpuhash_ring_t *q, unsigned timeout_ms)
{
	if (!q) return -1;
	for (int r = 0; r < q->rings; r++)
		if (q->next[r] > 1) { int rc = gpuhash_ring_wait(q, r, (long long)q->next[r] - 1, timeout_ms); if (rc) return rc; }
	return 0;
}

extern "C" int gpuhash_ring_ctas_per_ring(const gpuhash_ring_t *q) { return q ? q->ctas_per_ring : -1; }

/* globaltimer stamps (ns) of the latest batch of `ring`: [0] doorbell seen by the leader, [1] descriptor published to the
 * group, [2] last CTA finished its share, [3] completion mark written, [4] last CTA saw `go`, [5] leader started polling */
extern "C" int gpuhash_ring_trace(gpuhash_ring_t *q, int ring, unsigned long long out8[8])
{
	if (!q || ring < 0 || ring >= q->rings || !out8) return -1;
	cudaStream_t s;
	cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
	if (e != cudaSuccess) return (int)e;
	e = cudaMemcpyAsync(out8, q->trace_d + 8 * ring, sizeof(unsigned long long) * 8, cudaMemcpyDeviceToHost, s);
	if (e == cudaSuccess) e = cudaStreamSynchronize(s);
	cudaStreamDestroy(s);
	return (int)e;
}

/* K cycles through the rings (cycle i -> ring i % rings, the i-th batch of the pinned arrays, as gpuhash_bench_e2e).
 * There is no launch to bracket with events: total_ms is host wall clock from the first doorbell to the last completion
 * mark, i.e. it includes the host's own submit loop.  rtt_us (optional): median round trip of `rtt_reps` isolated
 * batches (doorbell -> completion mark seen by the host), the latency a lone batch pays. */
static int cmp_float(const void *a, const void *b) { float x = *(const float *)a, y = *(const float *)b; return x < y ? -1 : x > y; }

extern "C" int gpuhash_bench_ring(gpuhash_ring_t *q, const void *search_h, size_t n_search, void *out_h,
		const void *insert_h, size_t n_insert, int steps, gpuhash_bench_result_t *res, int rtt_reps, float *rtt_us)
{
	if (!q || !res || steps < 1) return -1;
	memset(res, 0, sizeof *res);
	int rc = gpuhash_ring_drain(q, 20000);
	if (rc) return rc;
	const double t0 = wall_ms();
	for (int i = 0; i < steps; i++) {
		long long t = gpuhash_ring_submit(q, i % q->rings,
				(const char *)search_h + (size_t)i * n_search * 8, n_search, (char *)out_h + (size_t)i * n_search * 8,
				NULL, 0, (const char *)insert_h + (size_t)i * n_insert * 12, n_insert);
		if (t < 0) return (int)t;
	}
	if ((rc = gpuhash_ring_drain(q, 60000)) != 0) return rc;
	res->total_ms = (float)(wall_ms() - t0);
	res->search_ops = (unsigned long long)steps * n_search;
	res->insert_ops = (unsigned long long)steps * n_insert;
	res->h2d_bytes = (unsigned long long)steps * (n_search * 8 + n_insert * 12);
	res->d2h_bytes = (unsigned long long)steps * n_search * 8;
	if (rtt_us && rtt_reps > 0) {
		float *lat = (float *)malloc(sizeof(float) * rtt_reps);
		if (!lat) return -1;
		for (int i = 0; i < rtt_reps; i++) {                              /* searches only: the same requests again, no table change */
			const int k = i % steps;
			const double a = wall_ms();
			long long t = gpuhash_ring_submit(q, 0, (const char *)search_h + (size_t)k * n_search * 8, n_search,
					(char *)out_h + (size_t)k * n_search * 8, NULL, 0, NULL, 0);
			if (t < 0 || (rc = gpuhash_ring_wait(q, 0, t, 20000)) != 0) { free(lat); return t < 0 ? (int)t : rc; }
			lat[i] = (float)((wall_ms() - a) * 1e3);
		}
		qsort(lat, rtt_reps, sizeof(float), cmp_float);
		*rtt_us = lat[rtt_reps / 2];
		free(lat);
	}
	return 0;
}
