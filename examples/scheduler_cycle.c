/*
 * scheduler_cycle.c -- a plain-C host driving the library the way Mega-KV's scheduler does
 * (src/mega_scheduler.c:306-532), in the three integration levels of INTEGRATION.md:
 *
 *   legacy   per worker and cycle: cudaMemcpyAsync H2D, gpu_hash_search / gpu_hash_delete / gpu_hash_insert on the
 *            worker's stream, cudaMemcpyAsync D2H, one cudaDeviceSynchronize per cycle -- the reference's own loop
 *            (:392-504) against the unchanged libgpuhash.h ABI
 *   submit   gpuhash_index_submit per worker (zero-copy from the pinned batch buffers) + gpuhash_index_sync
 *   ring     gpuhash_ring_submit per worker (one descriptor, no CUDA call) + gpuhash_ring_drain
 *
 * Workload: the LOCAL_TEST shape (src/mega_recv.c:634-768) with the survey's key stream -- every cycle each worker
 * SETs `nset` fresh keys and GETs `nget` keys drawn from everything SET in earlier cycles; a GET must come back with
 * the location that was SET (libgpuhash/test/insert_test.c:178-195 checks the same property).  Prints Mops/s and
 * exits non-zero on any wrong result.
 *
 *   gcc -O2 -std=gnu99 -Iinclude -I/usr/local/cuda/include examples/scheduler_cycle.c \
 *       -Lmegakv_b200/lib -lgpuhash -L/usr/local/cuda/lib64 -lcudart -lrt -lpthread -o examples/scheduler_cycle
 *   ./examples/scheduler_cycle [legacy|submit|ring] [workers=8] [cycles=50] [mem_p=30]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "libgpuhash.h"
#include "gpuhash_ex.h"

#define MAX_W 16
#define NGET 30000          /* config->batch_max_search_job is 32768 (src/mega.c:136) */
#define NSET 4096           /* batch_max_insert_job (src/mega.c:143) */

static uint64_t splitmix(uint64_t idx)
{
	uint64_t z = 1 + (idx + 1) * 0x9E3779B97F4A7C15ULL;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

static void key_to_req(uint64_t key, uint32_t *sig, uint32_t *hash)   /* src/mega_recv.c:361-362 */
{
	*hash = (uint32_t)(key >> 32);
	*sig = (uint32_t)key ? (uint32_t)key : 1u;                         /* 0 is the empty-slot marker */
}

static double now_s(void)
{
	struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

typedef struct {
	selem_t *search_in; loc_t *search_out; ielem_t *insert_in;        /* pinned, like mega_recv.c:154-156,176 */
	uint32_t *expect;                                                  /* host only: location each GET must return */
	selem_t *search_in_d; loc_t *search_out_d; ielem_t *insert_in_d;   /* legacy mode: device staging (mega_recv.c:132-150) */
	ielem_t **blk_ptr_d; int *blk_num_d;
	cudaStream_t stream;
	int nget, nset;
} worker_t;

int main(int argc, char **argv)
{
	const char *mode = argc > 1 ? argv[1] : "submit";
	int W = argc > 2 ? atoi(argv[2]) : 8, cycles = argc > 3 ? atoi(argv[3]) : 50, mem_p = argc > 4 ? atoi(argv[4]) : 30;
	if (W < 1 || W > MAX_W || cycles < 2) { fprintf(stderr, "bad arguments\n"); return 2; }
	const int legacy = !strcmp(mode, "legacy"), ring = !strcmp(mode, "ring");

	gpuhash_geom_t geom;
	if (gpuhash_geom_init(&geom, mem_p, GPUHASH_CUCKOO)) { fprintf(stderr, "bad mem_p\n"); return 2; }
	gpuhash_set_default_geom(&geom);                                   /* what the three legacy symbols use */
	gpuhash_index_t *ix = gpuhash_index_create(mem_p, GPUHASH_CUCKOO, W, NGET, NSET, 1);
	if (!ix) { fprintf(stderr, "gpuhash_index_create failed\n"); return 1; }
	gpuhash_index_set_zero_copy(ix, 1);
	bucket_t *table = (bucket_t *)gpuhash_index_table(ix);
	gpuhash_ring_t *q = ring ? gpuhash_ring_create(&geom, table, W, 2, 0, 0) : NULL;
	if (ring && !q) { fprintf(stderr, "gpuhash_ring_create failed\n"); return 1; }

	worker_t wk[MAX_W];
	memset(wk, 0, sizeof wk);
	for (int w = 0; w < W; w++) {
		CUDA_SAFE_CALL(cudaHostAlloc((void **)&wk[w].search_in, NGET * sizeof(selem_t), 0));
		CUDA_SAFE_CALL(cudaHostAlloc((void **)&wk[w].search_out, 2 * NGET * sizeof(loc_t), 0));
		CUDA_SAFE_CALL(cudaHostAlloc((void **)&wk[w].insert_in, NSET * sizeof(ielem_t), 0));
		wk[w].expect = (uint32_t *)malloc(NGET * sizeof(uint32_t));
		if (legacy) {
			CUDA_SAFE_CALL(cudaMalloc((void **)&wk[w].search_in_d, NGET * sizeof(selem_t)));
			CUDA_SAFE_CALL(cudaMalloc((void **)&wk[w].search_out_d, 2 * NGET * sizeof(loc_t)));
			CUDA_SAFE_CALL(cudaMalloc((void **)&wk[w].insert_in_d, NSET * sizeof(ielem_t)));
			CUDA_SAFE_CALL(cudaMalloc((void **)&wk[w].blk_ptr_d, sizeof(ielem_t *)));
			CUDA_SAFE_CALL(cudaMalloc((void **)&wk[w].blk_num_d, sizeof(int)));
			CUDA_SAFE_CALL(cudaMemcpy(wk[w].blk_ptr_d, &wk[w].insert_in_d, sizeof(ielem_t *), cudaMemcpyHostToDevice));
			CUDA_SAFE_CALL(cudaStreamCreate(&wk[w].stream));
		}
	}

	uint64_t next_key = 0, seed = 12345, wrong = 0, gets = 0, sets = 0;
	double t_gpu = 0;
	for (int c = 0; c < cycles; c++) {
		/* ---- the receivers' part: fill the batch buffers (not timed: that is CPU packet work in Mega-KV) ---- */
		const uint64_t known = next_key;                               /* keys SET in earlier cycles */
		for (int w = 0; w < W; w++) {
			wk[w].nset = NSET; wk[w].nget = known ? NGET : 0;
			for (int i = 0; i < wk[w].nset; i++) {
				const uint64_t k = next_key++;
				key_to_req(splitmix(k), &wk[w].insert_in[i].sig, &wk[w].insert_in[i].hash);
				wk[w].insert_in[i].loc = (uint32_t)(k + 1);
			}
			for (int i = 0; i < wk[w].nget; i++) {
				seed = seed * 6364136223846793005ULL + 1442695040888963407ULL;
				const uint64_t k = (seed >> 11) % known;
				key_to_req(splitmix(k), &wk[w].search_in[i].sig, &wk[w].search_in[i].hash);
				wk[w].expect[i] = (uint32_t)(k + 1);
			}
		}
		/* ---- the scheduler's part ---- */
		const double t0 = now_s();
		for (int w = 0; w < W; w++) {
			worker_t *b = &wk[w];
			if (legacy) {                                              /* mega_scheduler.c:392-504, verbatim in shape */
				if (b->nget) {
					CUDA_SAFE_CALL(cudaMemcpyAsync(b->search_in_d, b->search_in, b->nget * sizeof(selem_t), cudaMemcpyHostToDevice, b->stream));
					CUDA_SAFE_CALL(cudaMemsetAsync(b->search_out_d, 0, 2 * b->nget * sizeof(loc_t), b->stream));
					gpu_hash_search(b->search_in_d, b->search_out_d, table, b->nget, 24576, 256, b->stream);
					CUDA_SAFE_CALL(cudaMemcpyAsync(b->search_out, b->search_out_d, 2 * b->nget * sizeof(loc_t), cudaMemcpyDeviceToHost, b->stream));
				}
				CUDA_SAFE_CALL(cudaMemcpyAsync(b->insert_in_d, b->insert_in, b->nset * sizeof(ielem_t), cudaMemcpyHostToDevice, b->stream));
				CUDA_SAFE_CALL(cudaMemcpyAsync(b->blk_num_d, &b->nset, sizeof(int), cudaMemcpyHostToDevice, b->stream));
				gpu_hash_insert(table, b->blk_ptr_d, b->blk_num_d, 1, b->stream);
			} else if (ring) {
				if (gpuhash_ring_submit(q, w, b->search_in, b->nget, b->search_out, NULL, 0, b->insert_in, b->nset) < 0) { fprintf(stderr, "ring submit failed\n"); return 1; }
			} else {
				if (gpuhash_index_submit(ix, w, b->search_in, b->nget, b->search_out, NULL, 0, b->insert_in, b->nset)) { fprintf(stderr, "submit failed\n"); return 1; }
			}
		}
		if (ring) { if (gpuhash_ring_drain(q, 20000)) { fprintf(stderr, "ring drain failed\n"); return 1; } }
		else CUDA_SAFE_CALL(cudaDeviceSynchronize());                  /* mega_scheduler.c:504 */
		if (c > 0) t_gpu += now_s() - t0;                              /* the first cycle warms up */
		/* ---- the senders' part: take out[2i], else out[2i+1] (src/mega_send.c:411-414) ---- */
		for (int w = 0; w < W; w++) {
			for (int i = 0; i < wk[w].nget; i++) {
				loc_t loc = wk[w].search_out[2 * i];
				if (loc == 0) loc = wk[w].search_out[2 * i + 1];
				if (loc != wk[w].expect[i]) wrong++;
			}
			if (c > 0) { gets += wk[w].nget; sets += wk[w].nset; }
		}
	}
	printf("%s: %d workers x %d cycles, %llu GETs + %llu SETs in %.3f ms of scheduler time -> %.1f Mops/s, %llu wrong results\n",
			mode, W, cycles, (unsigned long long)gets, (unsigned long long)sets, t_gpu * 1e3, (gets + sets) / t_gpu / 1e6,
			(unsigned long long)wrong);
	if (q) gpuhash_ring_destroy(q);
	gpuhash_index_destroy(ix);
	return wrong ? 1 : 0;
}
