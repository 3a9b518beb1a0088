/*
 * libgpuhash.h -- the C ABI Mega-KV's scheduler links against.
 *
 * Same three entry points, argument order and argument meaning as the reference
 * (libgpuhash/libgpuhash.h:29-51), so src/mega_scheduler.c:409,452,496 and
 * libgpuhash/test/insert_test.c:145,173,207 compile and link unchanged.  The
 * implementation behind them is new sm_100a code (megakv_b200/csrc/).
 *
 * Contract kept from the reference:
 *   - every pointer is a DEVICE pointer owned by the caller; the table is one
 *     cudaMalloc(HT_SIZE) that the caller zero-filled (all-zero == empty);
 *   - calls are asynchronous launches on `stream` (0 == legacy default stream)
 *     and return nothing; a violated precondition aborts the process, a CUDA
 *     failure shows up at the caller's next CUDA_SAFE_CALL / synchronize;
 *   - `out` of gpu_hash_search holds 2*num_elem loc_t: [2i] is the hit in the
 *     key's first bucket, [2i+1] the hit in its alternate bucket, 0 == miss.
 *     (The reference only stores hits and relies on the caller's memset; this
 *     library stores both words of every query, misses as 0, so the memset is
 *     redundant but harmless.)
 *   - num_thread / threads_per_blk are launch-shape hints of the old kernels;
 *     they are accepted (any value the reference accepted, and the ones its
 *     16-bit rounding bug broke) and otherwise ignored.
 *   - the table's bytes belong to the library: slots are kept as {sig, loc}
 *     pairs inside each 64 B bucket unless the reference byte layout is selected
 *     (see gpu_hash.h); hosts that memcpy bucket_t images must select it;
 *   - geometry (MEM_P) and policy (HASH_CUCKOO / HASH_2CHOICE) default to the
 *     values of gpu_hash.h this header is compiled with by the LIBRARY; use
 *     gpuhash_set_default_geom() from gpuhash_ex.h to change them at run time.
 */
#ifndef _LIBGPUHASH_H_
#define _LIBGPUHASH_H_

#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime.h>
#include "gpu_hash.h"

#ifdef __cplusplus
extern "C" {
#endif

/* reference: libgpuhash.h:29-36, gpu_hash.cu:482-518 */
void gpu_hash_search(selem_t *in, loc_t *out, bucket_t *hash_table,
		int num_elem, int num_thread, int threads_per_blk,
		cudaStream_t stream);

/* reference: libgpuhash.h:38-43, gpu_hash.cu:521-556.
 * blk_input is a device array of num_blks device pointers, blk_elem_num a
 * device array of num_blks counts (read on the device, never on the host). */
void gpu_hash_insert(bucket_t *hash_table, ielem_t **blk_input,
		int *blk_elem_num, int num_blks, cudaStream_t stream);

/* reference: libgpuhash.h:45-51, gpu_hash.cu:558-593 */
void gpu_hash_delete(delem_t *in, bucket_t *hash_table,
		int num_elem, int num_thread, int threads_per_blk,
		cudaStream_t stream);

/* Declared by the reference (libgpuhash.h:53-62) but never defined there.
 * Here it is one launch that applies the delete batch and then the insert
 * batch with the same results as gpu_hash_delete followed by gpu_hash_insert
 * on the same stream. */
void gpu_delete_insert(bucket_t *hash_table, delem_t *delete_in,
		uint32_t num_delete_job, ielem_t **insert_blk_input,
		int *insert_blk_elem_num, int num_insert_blks,
		uint32_t num_delete_thread, uint32_t threads_per_blk,
		cudaStream_t stream);

#ifdef __cplusplus
}
#endif

/* reference: libgpuhash.h:64-70 (used throughout src/ and the tests) */
#define CUDA_SAFE_CALL(call) do {                                             \
	cudaError_t err = call;                                                   \
	if (cudaSuccess != err) {                                                 \
		fprintf(stderr, "Cuda error in file '%s' in line %i : %s.\n",         \
				__FILE__, __LINE__, cudaGetErrorString(err));                 \
		exit(EXIT_FAILURE);                                                   \
	} } while (0)

/* The reference's CUDA_SAFE_CALL_SYNC (libgpuhash.h:72-79) does not compile if
 * expanded; this one does what its name says. */
#define CUDA_SAFE_CALL_SYNC(call) do {                                        \
	CUDA_SAFE_CALL(call);                                                     \
	CUDA_SAFE_CALL(cudaDeviceSynchronize());                                  \
	} while (0)

#endif /* _LIBGPUHASH_H_ */
