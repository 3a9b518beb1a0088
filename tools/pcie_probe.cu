// pcie_probe.cu -- what can the host link of this box do, and does the placement of the pinned pages matter?
//
// The e2e leg of bench.py moves 8 B in + 8 B out per search over PCIe and measured ~21 GB/s per direction whatever
// the launch shape (staged, zero-copy, graph).  This probe separates the possible causes:
//   1. topology: NUMA nodes of the box, the GPU's node (/sys/bus/pci/devices/<bdf>/numa_node), CPUs allowed
//   2. large copies (64 MiB) H2D, D2H, both at once, from cudaHostAlloc memory and from mmap+mbind(node)+cudaHostRegister
//      memory on every node
//   3. the bench's shape: 498 KB H2D + 498 KB D2H per step round-robin over S streams (S = 1, 4, 32)
//   4. zero-copy kernels: streaming read of pinned host memory, streaming write to it, and both
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pcie_probe tools/pcie_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <unistd.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

static long sys_mbind(void *addr, unsigned long len, int mode, const unsigned long *mask, unsigned long maxnode, unsigned flags)
{
	return syscall(SYS_mbind, addr, len, mode, mask, maxnode, flags);
}

static int read_int(const char *path, int dflt)
{
	FILE *f = fopen(path, "r");
	if (!f) return dflt;
	int v = dflt;
	if (fscanf(f, "%d", &v) != 1) v = dflt;
	fclose(f);
	return v;
}

static void cat(const char *path)
{
	FILE *f = fopen(path, "r");
	if (!f) { printf("  %s: (absent)\n", path); return; }
	char buf[512];
	if (fgets(buf, sizeof buf, f)) { buf[strcspn(buf, "\n")] = 0; printf("  %s: %s\n", path, buf); }
	fclose(f);
}

// pinned memory whose pages sit on `node` (node < 0: wherever first touch puts them)
static void *alloc_on_node(size_t bytes, int node)
{
	void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
	if (p == MAP_FAILED) return NULL;
	if (node >= 0) {
		unsigned long mask[4] = {0, 0, 0, 0};
		mask[node / 64] |= 1UL << (node % 64);
		if (sys_mbind(p, bytes, 2 /* MPOL_BIND */, mask, 256, 0) != 0) printf("  mbind(node %d) failed: %s\n", node, strerror(errno));
	}
	memset(p, 1, bytes);
	if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { printf("  cudaHostRegister failed\n"); cudaGetLastError(); munmap(p, bytes); return NULL; }
	return p;
}

static float time_copies(void *d0, void *d1, void *h0, void *h1, size_t bytes, int reps, int dir /* 1 h2d, 2 d2h, 3 both */)
{
	cudaStream_t a, b; cudaEvent_t e0, e1, eb;
	CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&eb, cudaEventDisableTiming));
	CK(cudaDeviceSynchronize());
	CK(cudaEventRecord(e0, a)); CK(cudaStreamWaitEvent(b, e0, 0));
	for (int r = 0; r < reps; r++) {
		if (dir & 1) CK(cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, a));
		if (dir & 2) CK(cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, b));
	}
	CK(cudaEventRecord(eb, b)); CK(cudaStreamWaitEvent(a, eb, 0)); CK(cudaEventRecord(e1, a));
	CK(cudaEventSynchronize(e1));
	float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
	cudaStreamDestroy(a); cudaStreamDestroy(b); cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(eb);
	return ms;
}

__global__ void __launch_bounds__(256) zc_kernel(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n16, int mode)
{
	uint4 acc = make_uint4(0, 0, 0, 0);
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
		uint4 v = make_uint4((uint32_t)i, 1, 2, 3);
		if (mode & 1) v = in[i];
		if (mode & 2) out[i] = v;
		else { acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
	}
	if (!(mode & 2) && (acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x12345678u) out[0] = acc;
}

int main(int argc, char **argv)
{
	int dev = argc > 1 ? atoi(argv[1]) : 0;
	CK(cudaSetDevice(dev));
	cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, dev));
	char bdf[32]; CK(cudaDeviceGetPCIBusId(bdf, sizeof bdf, dev));
	for (char *c = bdf; *c; c++) if (*c >= 'A' && *c <= 'Z') *c += 32;
	printf("== topology\n  device %d %s bdf %s, asyncEngineCount %d\n", dev, pr.name, bdf, pr.asyncEngineCount);
	char path[256];
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bdf); cat(path);
	int gpu_node = read_int(path, -1);
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bdf); cat(path);
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/current_link_speed", bdf); cat(path);
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/current_link_width", bdf); cat(path);
	snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/max_link_speed", bdf); cat(path);
	cat("/sys/devices/system/node/online");
	cat("/sys/devices/system/node/has_cpu");
	cat("/sys/devices/system/node/has_memory");
	int nodes = 0;
	for (int n = 0; n < 16; n++) {
		snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", n);
		if (access(path, R_OK) == 0) { cat(path); nodes = n + 1; }
	}
	cpu_set_t cs; CPU_ZERO(&cs);
	if (sched_getaffinity(0, sizeof cs, &cs) == 0) {
		printf("  allowed cpus (%d):", CPU_COUNT(&cs));
		int first = -1;
		for (int c = 0; c <= 1024; c++) {
			bool in = c < 1024 && CPU_ISSET(c, &cs);
			if (in && first < 0) first = c;
			if (!in && first >= 0) { printf(" %d-%d", first, c - 1); first = -1; }
		}
		printf("\n  running on cpu %d\n", sched_getcpu());
	}
	printf("  gpu numa node %d, nodes seen %d\n", gpu_node, nodes);

	const size_t big = 64u << 20;
	void *d0, *d1; CK(cudaMalloc(&d0, big)); CK(cudaMalloc(&d1, big));
	printf("== large copies, %zu MiB x 8, GB/s per direction (h2d | d2h | both: h2d+d2h each)\n", big >> 20);
	for (int src = -2; src < nodes; src++) {
		void *h0, *h1;
		if (src == -2) { CK(cudaHostAlloc(&h0, big, cudaHostAllocDefault)); CK(cudaHostAlloc(&h1, big, cudaHostAllocDefault)); memset(h0, 1, big); memset(h1, 1, big); }
		else { h0 = alloc_on_node(big, src); h1 = alloc_on_node(big, src); if (!h0 || !h1) continue; }
		time_copies(d0, d1, h0, h1, big, 2, 3);
		float a = time_copies(d0, d1, h0, h1, big, 8, 1), b = time_copies(d0, d1, h0, h1, big, 8, 2), c = time_copies(d0, d1, h0, h1, big, 8, 3);
		const double gb = 8.0 * big / 1e9;
		printf("  %-28s %6.1f | %6.1f | %6.1f\n", src == -2 ? "cudaHostAlloc" : src == -1 ? "mmap first-touch + register" : (snprintf(path, sizeof path, "mmap mbind node %d + register", src), path),
			gb / (a / 1e3), gb / (b / 1e3), gb / (c / 1e3));
		// bench shape on this memory: per step 498 KB up, kernel-less, 498 KB down, over S streams
		const size_t step = 62259 * 8;
		for (int S = 1; S <= 32; S *= (S == 1 ? 4 : 8)) {
			cudaStream_t st[32]; cudaEvent_t e0, e1, done[32];
			for (int k = 0; k < S; k++) { CK(cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&done[k], cudaEventDisableTiming)); }
			CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
			const int steps = 128;
			for (int pass = 0; pass < 2; pass++) {
				CK(cudaDeviceSynchronize());
				CK(cudaEventRecord(e0, st[0]));
				for (int k = 1; k < S; k++) CK(cudaStreamWaitEvent(st[k], e0, 0));
				for (int i = 0; i < steps; i++) {
					cudaStream_t s = st[i % S];
					CK(cudaMemcpyAsync((char *)d0 + (size_t)(i % 128) * step, (char *)h0 + (size_t)(i % 128) * step, step, cudaMemcpyHostToDevice, s));
					CK(cudaMemcpyAsync((char *)h1 + (size_t)(i % 128) * step, (char *)d1 + (size_t)(i % 128) * step, step, cudaMemcpyDeviceToHost, s));
				}
				for (int k = 1; k < S; k++) { CK(cudaEventRecord(done[k], st[k])); CK(cudaStreamWaitEvent(st[0], done[k], 0)); }
				CK(cudaEventRecord(e1, st[0])); CK(cudaEventSynchronize(e1));
			}
			float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
			printf("      bench-shaped copies, %2d streams: %.2f us per step -> %.1f GB/s per direction\n", S, ms * 1e3 / steps, steps * (double)step / 1e9 / (ms / 1e3));
			for (int k = 0; k < S; k++) { cudaStreamDestroy(st[k]); cudaEventDestroy(done[k]); }
			cudaEventDestroy(e0); cudaEventDestroy(e1);
		}
		// the same bytes as ONE batch per direction and cycle (32 workers x 498 KB), each direction on its own stream:
		// plain cudaMemcpyAsync calls back to back, and cudaMemcpyBatchAsync
		for (int batch_api = 0; batch_api <= 1; batch_api++) {
			cudaStream_t up, dn; cudaEvent_t e0, e1, ed;
			CK(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&dn, cudaStreamNonBlocking));
			CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreateWithFlags(&ed, cudaEventDisableTiming));
			const int W = 32, cycles = 8;
			void *dsts_u[W], *srcs_u[W], *dsts_d[W], *srcs_d[W]; size_t sizes[W];
			for (int w = 0; w < W; w++) {
				dsts_u[w] = (char *)d0 + (size_t)w * step; srcs_u[w] = (char *)h0 + (size_t)w * step;
				dsts_d[w] = (char *)h1 + (size_t)w * step; srcs_d[w] = (char *)d1 + (size_t)w * step; sizes[w] = step;
			}
			float ms = 0; bool ok = true;
			for (int pass = 0; pass < 2 && ok; pass++) {
				CK(cudaDeviceSynchronize());
				CK(cudaEventRecord(e0, up)); CK(cudaStreamWaitEvent(dn, e0, 0));
				for (int c = 0; c < cycles && ok; c++) {
					if (batch_api) {
						cudaMemcpyAttributes at; memset(&at, 0, sizeof at); at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
						size_t idx0 = 0, fail = 0;
						cudaError_t e = cudaMemcpyBatchAsync(dsts_u, srcs_u, sizes, W, &at, &idx0, 1, &fail, up);
						if (e == cudaSuccess) e = cudaMemcpyBatchAsync(dsts_d, srcs_d, sizes, W, &at, &idx0, 1, &fail, dn);
						if (e != cudaSuccess) { printf("      cudaMemcpyBatchAsync: %s\n", cudaGetErrorString(e)); cudaGetLastError(); ok = false; }
					} else {
						for (int w = 0; w < W; w++) CK(cudaMemcpyAsync(dsts_u[w], srcs_u[w], step, cudaMemcpyHostToDevice, up));
						for (int w = 0; w < W; w++) CK(cudaMemcpyAsync(dsts_d[w], srcs_d[w], step, cudaMemcpyDeviceToHost, dn));
					}
				}
				CK(cudaEventRecord(ed, dn)); CK(cudaStreamWaitEvent(up, ed, 0)); CK(cudaEventRecord(e1, up)); CK(cudaEventSynchronize(e1));
				CK(cudaEventElapsedTime(&ms, e0, e1));
			}
			if (ok) printf("      32 x 498 KB per direction and cycle, one stream per direction, %s: %.1f GB/s per direction\n",
					batch_api ? "cudaMemcpyBatchAsync" : "cudaMemcpyAsync calls", cycles * W * (double)step / 1e9 / (ms / 1e3));
			cudaStreamDestroy(up); cudaStreamDestroy(dn); cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(ed);
		}
		// zero-copy kernels on this memory
		void *hd0, *hd1; CK(cudaHostGetDevicePointer(&hd0, h0, 0)); CK(cudaHostGetDevicePointer(&hd1, h1, 0));
		for (int mode = 1; mode <= 3; mode++) {
			for (int blocks = 148; blocks <= 148 * 16; blocks *= 4) {
				cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
				float ms = 0;
				for (int pass = 0; pass < 2; pass++) {
					CK(cudaEventRecord(e0));
					zc_kernel<<<blocks, 256>>>((const uint4 *)(mode & 1 ? hd0 : d0), (uint4 *)(mode & 2 ? hd1 : d1), big / 16, mode);
					CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
				}
				printf("      zero-copy kernel %-10s %4d CTAs: %.1f GB/s per direction\n", mode == 1 ? "read" : mode == 2 ? "write" : "read+write", blocks, big / 1e9 / (ms / 1e3));
				cudaEventDestroy(e0); cudaEventDestroy(e1);
			}
		}
		if (src == -2) { cudaFreeHost(h0); cudaFreeHost(h1); }
		else { cudaHostUnregister(h0); cudaHostUnregister(h1); munmap(h0, big); munmap(h1, big); }
	}
	return 0;
}
