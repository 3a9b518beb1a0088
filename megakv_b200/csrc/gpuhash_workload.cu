/*
 * gpuhash_workload.cu -- synthetic request streams generated on the device (bench tooling).
 *
 * Same key stream as megakv_b200/keystream.py and oracle/gpuhash_oracle.c:orc_keys_fill
 * (SURVEY.md 8(d)): key_i = i-th splitmix64 output from state `seed`; hash = high 32 bits,
 * sig = low 32 bits with 0 -> 1 (src/mega_recv.c:350,361-362), loc = i + 1.
 * The reference builds its benchmark inputs on the host with rand() (libgpuhash/test/back/
 * py_search_stream.c:128-129); at 2^29 preloaded keys that would be minutes of host time and
 * 6 GiB over PCIe, so the generators run where the data is needed.
 */
#include <stdint.h>
#include <cuda_runtime.h>
#include "gpuhash_ex.h"

namespace {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

__device__ __forceinline__ uint64_t key_at(uint64_t seed, uint64_t idx)
{
	return mix64(seed + (idx + 1) * 0x9E3779B97F4A7C15ULL);
}

__device__ __forceinline__ void key_to_req(uint64_t key, uint32_t &sig, uint32_t &hash)
{
	sig = (uint32_t)key; hash = (uint32_t)(key >> 32);
	if (sig == 0) sig = 1;
}

__global__ void gen_inserts_kernel(uint32_t *iel, uint32_t *sel, uint64_t seed, uint64_t first, size_t n)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint32_t sig, hash; key_to_req(key_at(seed, first + i), sig, hash);
		if (iel) { iel[3 * i] = sig; iel[3 * i + 1] = hash; iel[3 * i + 2] = (uint32_t)(first + i + 1); }
		if (sel) { sel[2 * i] = sig; sel[2 * i + 1] = hash; }
	}
}

/* theta == 0: uniform index in [0, population).  0 < theta < 1: Zipf rank by Gray et al. (the method of
 * src/zipf.h:137-183, exact pow), rank 0 = most popular = key 0. */
__global__ void gen_queries_kernel(uint32_t *sel, uint32_t *expect_loc, uint64_t seed, uint64_t population,
		size_t n, uint64_t rng_seed, double theta, double zetan, int triples)
{
	/* constants of Gray's method, computed on the device so the host side needs no libm
	 * (the reference's link line has none: libgpuhash/test/Makefile:5) */
	double eta = 0, alpha = 0, thres = 0;
	if (theta > 0.0) {
		double zeta2 = 1.0 + pow(0.5, theta);
		alpha = 1.0 / (1.0 - theta);
		eta = (1.0 - pow(2.0 / (double)population, 1.0 - theta)) / (1.0 - zeta2 / zetan);
		thres = zeta2;
	}
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint64_t r = mix64(rng_seed + i * 0xD1342543DE82EF95ULL);
		uint64_t idx;
		if (theta == 0.0) {
			idx = (uint64_t)(((unsigned __int128)r * population) >> 64);
		} else {
			double u = (double)(r >> 11) * (1.0 / 9007199254740992.0);
			double uz = u * zetan;
			if (uz < 1.0) idx = 0;
			else if (uz < thres) idx = 1;
			else idx = (uint64_t)((double)population * pow(eta * (u - 1.0) + 1.0, alpha));
			if (idx >= population) idx = population - 1;
		}
		uint32_t sig, hash; key_to_req(key_at(seed, idx), sig, hash);
		if (triples) { sel[3 * i] = sig; sel[3 * i + 1] = hash; sel[3 * i + 2] = (uint32_t)(idx + 1); }
		else { sel[2 * i] = sig; sel[2 * i + 1] = hash; }
		if (expect_loc) expect_loc[i] = (uint32_t)(idx + 1);
	}
}

/* ---- the reference's own rank generator, src/zipf.h:44-183, bit for bit (BASELINE configs[2] is quoted on it) ----
 * Every double operation is an explicit round-to-nearest intrinsic: nvcc would otherwise contract a * b + c into one FMA,
 * which rounds once where the host code the reference compiles to rounds twice. */
__device__ __forceinline__ double ref_pow_approx(double a, double b)                  /* zipf.h:44-71 */
{
	int whole = (int)b;
	int hi = __double2hiint(a);
	hi = (int)__dadd_rn(__dmul_rn(__dsub_rn(b, (double)whole), (double)(hi - 1072632447)), 1072632447.);
	const double frac = __hiloint2double(hi, 0);
	double r = 1.;
	for (; whole; whole >>= 1, a = __dmul_rn(a, a))
		if (whole & 1) r = __dmul_rn(r, a);
	return __dmul_rn(r, frac);
}

/* state i + 1 of the 48-bit LCG x <- a x + c (zipf.h:117-126) started at x0: the affine map raised to the power i + 1 by
 * squaring, so that every thread draws its own element of the SAME sequence the sequential generator walks */
__device__ __forceinline__ uint64_t ref_lcg_at(uint64_t x0, uint64_t steps)
{
	uint64_t ra = 1, rc = 0, a = 0x5deece66dULL, c = 0xbULL;
	for (; steps; steps >>= 1) {
		if (steps & 1) { ra = ra * a; rc = rc * a + c; }
		c = c * a + c; a = a * a;
	}
	return (ra * x0 + rc) & ((1ULL << 48) - 1);
}

__global__ void gen_queries_ref_kernel(uint32_t *sel, uint32_t *expect_loc, uint64_t seed, uint64_t population,
		size_t n, uint64_t rand_seed, uint64_t first, double theta, double zetan, int triples)
{
	double eta = 0, alpha = 0, thres = 0;
	const double dbl_n = (double)population;
	if (theta > 0.0) {                                                               /* zipf.h:94-97, 143-146 */
		alpha = __ddiv_rn(1., __dsub_rn(1., theta));
		thres = __dadd_rn(1., ref_pow_approx(0.5, theta));
		const double zeta2 = __dadd_rn(__ddiv_rn(1., ref_pow_approx(1., theta)), __ddiv_rn(1., ref_pow_approx(2., theta)));
		eta = __ddiv_rn(__dsub_rn(1., ref_pow_approx(__ddiv_rn(2., dbl_n), __dsub_rn(1., theta))), __dsub_rn(1., __ddiv_rn(zeta2, zetan)));
	}
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint64_t x = ref_lcg_at(rand_seed, first + i + 1);
		const double u = __ddiv_rn((double)x, (double)((1ULL << 48) - 1));
		uint64_t idx;
		if (theta == 0.0) idx = (uint64_t)__dmul_rn(dbl_n, u);                          /* :161-165 */
		else {
			const double uz = __dmul_rn(u, zetan);
			if (uz < 1.0) idx = 0;
			else if (uz < thres) idx = 1;
			else idx = (uint64_t)__dmul_rn(dbl_n, ref_pow_approx(__dadd_rn(__dmul_rn(eta, __dsub_rn(u, 1.)), 1.), alpha));   /* :171-181 */
		}
		if (idx >= population) idx = population - 1;                                 /* (the reference would index past its key array) */
		uint32_t sig, hash; key_to_req(key_at(seed, idx), sig, hash);
		if (triples) { sel[3 * i] = sig; sel[3 * i + 1] = hash; sel[3 * i + 2] = (uint32_t)(idx + 1); }
		else { sel[2 * i] = sig; sel[2 * i + 1] = hash; }
		if (expect_loc) expect_loc[i] = (uint32_t)(idx + 1);
	}
}

}  // namespace

extern "C" int gpuhash_gen_inserts(void *ielem_d, void *selem_d, uint64_t seed, uint64_t first, size_t n, void *stream)
{
	if (n == 0) return 0;
	size_t blocks = (n + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32;
	gen_inserts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint32_t *)ielem_d, (uint32_t *)selem_d, seed, first, n);
	return (int)cudaGetLastError();
}

extern "C" int gpuhash_gen_queries(void *selem_d, void *expect_loc_d, uint64_t seed, uint64_t population, size_t n,
		uint64_t rng_seed, double theta, double zetan, void *stream)
{
	if (n == 0) return 0;
	if (!selem_d || population < 2 || theta < 0.0 || theta >= 1.0) return -1;
	size_t blocks = (n + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32;
	gen_queries_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint32_t *)selem_d, (uint32_t *)expect_loc_d,
			seed, population, n, rng_seed, theta, zetan, 0);
	return (int)cudaGetLastError();
}

/* the same draw as gpuhash_gen_queries, emitted as (sig, hash, loc = key index + 1) triples: insert / delete requests
 * for keys of the population (uniform or Zipf) -- updates of present keys, deletes that repeat hot keys */
extern "C" int gpuhash_gen_requests(void *ielem_d, uint64_t seed, uint64_t population, size_t n,
		uint64_t rng_seed, double theta, double zetan, void *stream)
{
	if (n == 0) return 0;
	if (!ielem_d || population < 2 || theta < 0.0 || theta >= 1.0) return -1;
	size_t blocks = (n + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32;
	gen_queries_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint32_t *)ielem_d, nullptr,
			seed, population, n, rng_seed, theta, zetan, 1);
	return (int)cudaGetLastError();
}

/* Requests for keys whose ranks come from the REFERENCE's generator (src/zipf.h: mehcached_zipf_next, approximate pow,
 * 48-bit LCG seeded with rand_seed < 2^48), elements first .. first+n-1 of its sequence; theta in [0, 1), zetan as
 * mehcached_zeta(n, theta) computes it (megakv_b200.keystream.ref_zetan).  triples != 0: (sig, hash, loc = rank + 1) records. */
extern "C" int gpuhash_gen_requests_ref_zipf(void *out_d, void *expect_loc_d, uint64_t seed, uint64_t population, size_t n,
		uint64_t rand_seed, uint64_t first, double theta, double zetan, int triples, void *stream)
{
	if (n == 0) return 0;
	if (!out_d || population < 2 || theta < 0.0 || theta >= 1.0 || rand_seed >= (1ULL << 48) || (theta > 0.0 && !(zetan > 0.0))) return -1;
	size_t blocks = (n + 255) / 256; if (blocks > 148 * 32) blocks = 148 * 32;
	gen_queries_ref_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((uint32_t *)out_d, (uint32_t *)expect_loc_d,
			seed, population, n, rand_seed, first, theta, zetan, triples);
	return (int)cudaGetLastError();
}
