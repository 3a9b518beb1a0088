// atomic_probe.cu -- how many atomicAdd per second does ONE address take on B200?  (sizes the ticket counter of
// cycle_multi_kernel: one ticket per 64-request tile is ~0.35 G tickets/s at 22 Gops/s)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/atomic_probe tools/atomic_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

// mode 0: lane 0 of every warp, dependent chain (the next atomic needs the previous result) -- latency under contention
// mode 1: lane 0 of every warp, independent atomics (results summed) -- throughput
// mode 2: as 1 but with a fence before each (the publish pattern)
template <int MODE>
__global__ void probe(uint32_t *ctr, int addrs, int stride_words, int iters, uint32_t *sink)
{
	const unsigned lane = threadIdx.x & 31u;
	const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	uint32_t acc = 0;
	if (lane == 0) {
		uint32_t *p = ctr + (size_t)(warp % addrs) * stride_words;
		for (int i = 0; i < iters; i++) {
			if (MODE == 2) __threadfence();
			uint32_t v = atomicAdd(p, MODE == 0 ? (acc & 1u) + 1u : 1u);
			acc += v;
		}
	}
	if (acc == 0xdeadbeefu) *sink = acc;
}

int main()
{
	uint32_t *ctr, *sink;
	CK(cudaMalloc(&ctr, 1 << 20)); CK(cudaMemset(ctr, 0, 1 << 20)); CK(cudaMalloc(&sink, 4));
	cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
	const int iters = 64;
	for (int mode = 0; mode < 3; mode++)
		for (int warps_per_sm : {8, 24, 32})
			for (int addrs : {1, 2, 8, 64}) {
				const int blocks = 148 * warps_per_sm / 8;
				float best = 1e30f;
				for (int it = 0; it < 3; it++) {
					CK(cudaEventRecord(a));
					if (mode == 0) probe<0><<<blocks, 256>>>(ctr, addrs, 64, iters, sink);
					else if (mode == 1) probe<1><<<blocks, 256>>>(ctr, addrs, 64, iters, sink);
					else probe<2><<<blocks, 256>>>(ctr, addrs, 64, iters, sink);
					CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
					float ms; CK(cudaEventElapsedTime(&ms, a, b));
					if (it && ms < best) best = ms;
				}
				const double n = (double)blocks * 8 * iters;
				printf("{\"mode\": %d, \"warps_per_sm\": %d, \"addresses\": %d, \"atomics\": %.0f, \"us\": %.1f, \"G_per_s\": %.3f, \"ns_per_atomic_per_address\": %.2f}\n",
				       mode, warps_per_sm, addrs, n, best * 1e3, n / (best * 1e-3) / 1e9, best * 1e6 / (n / addrs));
			}
	return 0;
}
