// gather_flavours.cu -- which load instruction makes a random 32 B probe cost one DRAM sector instead of a 128 B line?
//
// Round-1 finding (profiles/r01_search_ncu.md): every random 32 B request of the search kernel shows up as 4 sectors
// at L2 (lts__t_sectors_srcunit_tex_op_read) and 4 sectors at DRAM (dram__sectors_read), whatever
// cudaLimitMaxL2FetchGranularity says.  This probe issues the same random-sector gather with different load
// flavours; run it plain for rates and under `ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum`
// for sectors per access.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/gather_flavours tools/gather_flavours.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

#define LD8(NAME, INSTR) \
__device__ __forceinline__ uint32_t NAME(const uint32_t *p) { uint32_t a, b, c, d, e, f, g, h; \
	asm volatile(INSTR " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
		: "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p) : "memory"); \
	return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h; }
#define LD4(NAME, INSTR) \
__device__ __forceinline__ uint32_t NAME(const uint32_t *p) { uint32_t a, b, c, d; \
	asm volatile(INSTR " {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory"); \
	return a ^ b ^ c ^ d; }
#define LD1(NAME, INSTR) \
__device__ __forceinline__ uint32_t NAME(const uint32_t *p) { uint32_t a; \
	asm volatile(INSTR " %0, [%1];" : "=r"(a) : "l"(p) : "memory"); return a; }

LD8(ld8_ca, "ld.global.ca.v8.u32")
LD8(ld8_na, "ld.global.L1::no_allocate.v8.u32")
LD8(ld8_cg, "ld.global.cg.v8.u32")
LD8(ld8_nc, "ld.global.nc.v8.u32")
LD8(ld8_cv, "ld.volatile.global.v8.u32")
LD8(ld8_ef, "ld.global.L1::no_allocate.L2::evict_first.v8.u32")
LD8(ld8_el, "ld.global.L1::no_allocate.L2::evict_last.v8.u32")
LD8(ld8_cs, "ld.global.cs.v8.u32")
LD8(ld8_lu, "ld.global.lu.v8.u32")
LD8(ld8_p64, "ld.global.L1::no_allocate.L2::64B.v8.u32")
LD8(ld8_p128, "ld.global.L1::no_allocate.L2::128B.v8.u32")
LD8(ld8_p256, "ld.global.L1::no_allocate.L2::256B.v8.u32")
LD8(ld8_ca64, "ld.global.ca.L2::64B.v8.u32")
LD8(ld8_rgpu, "ld.relaxed.gpu.global.v8.u32")
LD8(ld8_rsys, "ld.relaxed.sys.global.v8.u32")
LD8(ld8_acq, "ld.acquire.gpu.global.v8.u32")
LD4(ld4_p64, "ld.global.L1::no_allocate.L2::64B.v4.u32")
LD1(ld1_p64, "ld.global.L1::no_allocate.L2::64B.u32")
LD1(ld1_mmio, "ld.mmio.relaxed.sys.global.u32")
__device__ __forceinline__ uint32_t ld8_hint(const uint32_t *p, uint64_t pol) { uint32_t a, b, c, d, e, f, g, h;
	asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
		: "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p), "l"(pol) : "memory");
	return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h; }
__device__ __forceinline__ uint32_t ld8_hint64(const uint32_t *p, uint64_t pol) { uint32_t a, b, c, d, e, f, g, h;
	asm volatile("ld.global.L1::no_allocate.L2::cache_hint.L2::64B.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
		: "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p), "l"(pol) : "memory");
	return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h; }
LD4(ld4_ca, "ld.global.ca.v4.u32")
LD4(ld4_na, "ld.global.L1::no_allocate.v4.u32")
LD4(ld4_cg, "ld.global.cg.v4.u32")
LD1(ld1_ca, "ld.global.ca.u32")
LD1(ld1_cg, "ld.global.cg.u32")

enum { F_CA8, F_NA8, F_CG8, F_NC8, F_CV8, F_EF8, F_EL8, F_CS8, F_LU8, F_CA4x2, F_NA4x2, F_CG4x2, F_CA4, F_CG4, F_CA1, F_CG1,
       F_ATOM, F_CPASYNC16, F_BULK32, F_COOP8, F_PAIR2_1I, F_PAIR2_2I, F_QUAD4_1I,
       F_P64, F_P128, F_P256, F_CA64, F_RGPU, F_RSYS, F_ACQ, F_P64_16, F_P64_4, F_MMIO4, F_POL_EF, F_POL_EF64, F_POL_FRAC, F_POL_NOALLOC,
       F_PREF_L2, F_PREF_L2_EL, F_PAIR2_1I_64, F_RED, F_ATOM64, F_BULK64_HINT, F_COUNT };
static const char *kNames[F_COUNT] = {
	"ld.ca.v8 (32B)", "ld.L1::no_allocate.v8 (32B)", "ld.cg.v8 (32B)", "ld.nc.v8 (32B)", "ld.volatile.v8 (32B)",
	"ld.na.L2::evict_first.v8", "ld.na.L2::evict_last.v8", "ld.cs.v8 (32B)", "ld.lu.v8 (32B)",
	"2 x ld.ca.v4 (32B)", "2 x ld.na.v4 (32B)", "2 x ld.cg.v4 (32B)", "ld.ca.v4 (16B only)", "ld.cg.v4 (16B only)",
	"ld.ca.u32 (4B only)", "ld.cg.u32 (4B only)", "atom.add.u32 +0 (4B)", "cp.async.cg 2x16B -> smem", "cp.async.bulk 32B -> smem",
	"8 lanes x ld.ca.u32 (32B coop)",
	"64B bucket: 2 lanes, ONE ld.v8 instruction (n counts buckets)", "64B bucket: 1 thread, TWO ld.v8 instructions (n counts buckets)",
	"128B line: 4 lanes, ONE ld.v8 instruction (n counts lines)",
	"ld.na.L2::64B.v8 (32B)", "ld.na.L2::128B.v8 (32B)", "ld.na.L2::256B.v8 (32B)", "ld.ca.L2::64B.v8 (32B)",
	"ld.relaxed.gpu.v8 (32B)", "ld.relaxed.sys.v8 (32B)", "ld.acquire.gpu.v8 (32B)", "ld.na.L2::64B.v4 (16B only)", "ld.na.L2::64B.u32 (4B only)",
	"ld.mmio.relaxed.sys.u32 (4B only)", "ld.na + createpolicy evict_first 1.0", "ld.na.L2::64B + createpolicy evict_first 1.0",
	"ld.na + createpolicy fractional evict_last 0.125", "ld.na + createpolicy L2::evict_first/unchanged 0.5",
	"prefetch.global.L2 then ld.na.v8", "prefetch.global.L2::evict_last then ld.na.v8",
	"64B bucket: 2 lanes, ONE ld.na.L2::64B.v8 (n counts buckets)", "red.add.u32 +0 (4B)", "atom.add.u64 +0 (8B)",
	"cp.async.bulk 64B + L2::cache_hint evict_first -> smem (n counts 64B)" };

template <int F, int ILP>
__global__ void __launch_bounds__(256)
gather(uint32_t *table, uint64_t mask, size_t n, uint32_t seed, uint32_t *sink)
{
	__shared__ __align__(128) uint32_t stage[256 * 8];
	__shared__ __align__(8) uint64_t bar;
	uint32_t acc = 0;
	const size_t stride = (size_t)gridDim.x * blockDim.x * ILP;
	if (F == F_BULK32 || F == F_BULK64_HINT) {
		if (threadIdx.x == 0) {
			uint32_t a = (uint32_t)__cvta_generic_to_shared(&bar);
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(a), "r"(1));
		}
		__syncthreads();
	}
	uint64_t pol = 0;
	if (F == F_POL_EF || F == F_POL_EF64 || F == F_BULK64_HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	if (F == F_POL_FRAC) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 0.125;" : "=l"(pol));
	if (F == F_POL_NOALLOC) asm volatile("createpolicy.fractional.L2::evict_first.L2::evict_unchanged.b64 %0, 0.5;" : "=l"(pol));
	uint32_t phase = 0;
	for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * ILP; i < n; i += stride) {
		uint32_t v[ILP];
#pragma unroll
		for (int k = 0; k < ILP; k++) {
			size_t idx = i + k;
			if (F == F_COOP8) idx = (i / ILP / 8) * ILP + k + (size_t)seed;     // 8 neighbouring lanes share one sector
			if (F == F_PAIR2_1I) idx = (i / ILP / 2) * ILP + k;                 // 2 neighbouring lanes share one bucket
			if (F == F_QUAD4_1I) idx = (i / ILP / 4) * ILP + k;                 // 4 neighbouring lanes share one line
			if (F == F_PAIR2_1I_64) idx = (i / ILP / 2) * ILP + k;
			uint64_t u = mix64((uint64_t)idx * 0x9E3779B97F4A7C15ULL + seed) & mask;
			uint32_t *p = table + u * 8;
			if (F == F_CA8) v[k] = ld8_ca(p);
			else if (F == F_NA8) v[k] = ld8_na(p);
			else if (F == F_CG8) v[k] = ld8_cg(p);
			else if (F == F_NC8) v[k] = ld8_nc(p);
			else if (F == F_CV8) v[k] = ld8_cv(p);
			else if (F == F_EF8) v[k] = ld8_ef(p);
			else if (F == F_EL8) v[k] = ld8_el(p);
			else if (F == F_CS8) v[k] = ld8_cs(p);
			else if (F == F_LU8) v[k] = ld8_lu(p);
			else if (F == F_CA4x2) v[k] = ld4_ca(p) ^ ld4_ca(p + 4);
			else if (F == F_NA4x2) v[k] = ld4_na(p) ^ ld4_na(p + 4);
			else if (F == F_CG4x2) v[k] = ld4_cg(p) ^ ld4_cg(p + 4);
			else if (F == F_CA4) v[k] = ld4_ca(p);
			else if (F == F_CG4) v[k] = ld4_cg(p);
			else if (F == F_CA1) v[k] = ld1_ca(p);
			else if (F == F_CG1) v[k] = ld1_cg(p);
			else if (F == F_ATOM) v[k] = atomicAdd(p, 0u);
			else if (F == F_COOP8) v[k] = ld1_ca(p + (threadIdx.x & 7));
			else if (F == F_PAIR2_1I) v[k] = ld8_na(table + (u & ~1ULL) * 8 + (threadIdx.x & 1) * 8);
			else if (F == F_PAIR2_2I) { uint32_t *b = table + (u & ~1ULL) * 8; v[k] = ld8_na(b) ^ ld8_na(b + 8); }
			else if (F == F_QUAD4_1I) v[k] = ld8_na(table + (u & ~3ULL) * 8 + (threadIdx.x & 3) * 8);
			else if (F == F_P64) v[k] = ld8_p64(p);
			else if (F == F_P128) v[k] = ld8_p128(p);
			else if (F == F_P256) v[k] = ld8_p256(p);
			else if (F == F_CA64) v[k] = ld8_ca64(p);
			else if (F == F_RGPU) v[k] = ld8_rgpu(p);
			else if (F == F_RSYS) v[k] = ld8_rsys(p);
			else if (F == F_ACQ) v[k] = ld8_acq(p);
			else if (F == F_P64_16) v[k] = ld4_p64(p);
			else if (F == F_P64_4) v[k] = ld1_p64(p);
			else if (F == F_MMIO4) v[k] = ld1_mmio(p);
			else if (F == F_POL_EF || F == F_POL_FRAC || F == F_POL_NOALLOC) v[k] = ld8_hint(p, pol);
			else if (F == F_POL_EF64) v[k] = ld8_hint64(p, pol);
			else if (F == F_PREF_L2) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p) : "memory"); v[k] = 0; }
			else if (F == F_PREF_L2_EL) { asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p) : "memory"); v[k] = 0; }
			else if (F == F_PAIR2_1I_64) v[k] = ld8_p64(table + (u & ~1ULL) * 8 + (threadIdx.x & 1) * 8);
			else if (F == F_RED) { asm volatile("red.global.add.u32 [%0], 0;" :: "l"(p) : "memory"); v[k] = 0; }
			else if (F == F_ATOM64) v[k] = (uint32_t)atomicAdd((unsigned long long *)p, 0ULL);
			else if (F == F_BULK64_HINT) {
				uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[(threadIdx.x & 127) * 16]);
				uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
				if (threadIdx.x < 128)
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], 64, [%2], %3;"
						:: "r"(s), "l"(table + (u & ~1ULL) * 8), "r"(b), "l"(pol) : "memory");
				v[k] = 0;
			}
			else if (F == F_CPASYNC16) {
				uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[threadIdx.x * 8]);
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(p) : "memory");
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s + 16), "l"(p + 4) : "memory");
				v[k] = 0;
			} else if (F == F_BULK32) {
				uint32_t s = (uint32_t)__cvta_generic_to_shared(&stage[threadIdx.x * 8]);
				uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];"
					:: "r"(s), "l"(p), "r"(b) : "memory");
				v[k] = 0;
			}
		}
		if (F == F_PREF_L2 || F == F_PREF_L2_EL) {                           // the demand loads follow their prefetches
#pragma unroll
			for (int k = 0; k < ILP; k++) {
				uint64_t u = mix64((uint64_t)(i + k) * 0x9E3779B97F4A7C15ULL + seed) & mask;
				acc ^= ld8_na(table + u * 8);
			}
		} else if (F == F_BULK64_HINT) {
			__syncthreads();
			uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
			if (threadIdx.x == 0)
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(128 * ILP * 64) : "memory");
			uint32_t done = 0;
			while (!done)
				asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
					: "=r"(done) : "r"(b), "r"(phase) : "memory");
			phase ^= 1;
			acc ^= stage[threadIdx.x * 8 + 3];
			__syncthreads();
		} else if (F == F_CPASYNC16) {
			asm volatile("cp.async.commit_group;" ::: "memory");
			asm volatile("cp.async.wait_group 0;" ::: "memory");
			acc ^= stage[threadIdx.x * 8 + 3];
		} else if (F == F_BULK32) {
			// every thread issued ILP x 32 B; thread 0 posts the expected byte count for the whole CTA, all wait
			__syncthreads();
			uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
			if (threadIdx.x == 0)
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(256 * ILP * 32) : "memory");
			uint32_t done = 0;
			while (!done)
				asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
					: "=r"(done) : "r"(b), "r"(phase) : "memory");
			phase ^= 1;
			acc ^= stage[threadIdx.x * 8 + 3];
			__syncthreads();
		} else {
#pragma unroll
			for (int k = 0; k < ILP; k++) acc ^= v[k];
		}
	}
	if (acc == 0xDEADBEEFu && seed == 0x12345u) *sink = acc;
}

template <int F>
float run(uint32_t *table, uint64_t mask, size_t n, uint32_t seed, uint32_t *sink)
{
	constexpr int ILP = 4;
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	size_t blocks = (n + 256 * ILP - 1) / (256 * ILP);
	if ((F == F_BULK32 || F == F_BULK64_HINT) && blocks > 148 * 8) blocks = 148 * 8;              // CTA-wide barrier per round: keep CTAs resident
	float best = 1e30f;
	for (int it = 0; it < 3; it++) {
		CK(cudaEventRecord(a));
		gather<F, ILP><<<(unsigned)blocks, 256>>>(table, mask, n, seed + it, sink);
		CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
		CK(cudaGetLastError());
		float ms; CK(cudaEventElapsedTime(&ms, a, b));
		if (it && ms < best) best = ms;
	}
	return best;
}

int main(int argc, char **argv)
{
	int log2_bytes = argc > 1 ? atoi(argv[1]) : 34;
	size_t n = argc > 2 ? (size_t)atoll(argv[2]) : ((size_t)1 << 25);
	size_t bytes = (size_t)1 << log2_bytes;
	uint32_t *table, *sink;
	CK(cudaMalloc(&table, bytes)); CK(cudaMemset(table, 0, bytes)); CK(cudaMalloc(&sink, 4));
	uint64_t mask = bytes / 32 - 1;
	size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
	printf("{\"exp\": \"flavours\", \"table_log2\": %d, \"n\": %zu, \"l2_fetch_granularity\": %zu}\n", log2_bytes, n, gran);
	float ms;
#define RUN(F) ms = run<F>(table, mask, n, 77, sink); \
	printf("{\"flavour\": \"%s\", \"ms\": %.4f, \"Gaccess_per_s\": %.2f}\n", kNames[F], ms, n / (ms * 1e-3) / 1e9); fflush(stdout);
	RUN(F_CA8) RUN(F_NA8) RUN(F_CG8) RUN(F_NC8) RUN(F_CV8) RUN(F_EF8) RUN(F_EL8) RUN(F_CS8) RUN(F_LU8)
	RUN(F_CA4x2) RUN(F_NA4x2) RUN(F_CG4x2) RUN(F_CA4) RUN(F_CG4) RUN(F_CA1) RUN(F_CG1) RUN(F_ATOM) RUN(F_CPASYNC16) RUN(F_COOP8) RUN(F_PAIR2_1I) RUN(F_PAIR2_2I) RUN(F_QUAD4_1I) RUN(F_BULK32)
	RUN(F_P64) RUN(F_P128) RUN(F_P256) RUN(F_CA64) RUN(F_RGPU) RUN(F_RSYS) RUN(F_ACQ) RUN(F_P64_16) RUN(F_P64_4) RUN(F_MMIO4)
	RUN(F_POL_EF) RUN(F_POL_EF64) RUN(F_POL_FRAC) RUN(F_POL_NOALLOC) RUN(F_PREF_L2) RUN(F_PREF_L2_EL) RUN(F_PAIR2_1I_64) RUN(F_RED) RUN(F_ATOM64) RUN(F_BULK64_HINT)
	return 0;
}
