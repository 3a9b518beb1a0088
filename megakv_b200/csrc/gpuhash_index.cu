/*
 * gpuhash_index.cu -- the device side of Mega-KV's scheduler as a C object.
 *
 * Replaces, behind a host-buffer API, what src/mega_scheduler.c:392-504 does every cycle for
 * every worker (H2D of the worker's batch, the three launches in the order search -> delete ->
 * insert on the worker's stream, D2H of the search results, one synchronisation at the end of
 * the cycle), together with the per-worker device staging buffers of src/mega_recv.c:132-162.
 *
 * Differences that matter for speed, none for results:
 *   - the cudaMemsetAsync of search_out (mega_scheduler.c:406) is gone: the search kernel
 *     writes both result words of every request;
 *   - the 8 per-segment H2D copies + 1 count copy of the insert path (mega_scheduler.c:484-494)
 *     are one copy: the host-buffer API takes the insert batch as one array and the flat insert
 *     kernel is sized on the host;
 *   - streams are non-blocking, so workers overlap H2D, kernels and D2H freely.
 *
 * Plain C-style code, no libstdc++ (see libgpuhash.cu).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime.h>

#include "gpuhash_ex.h"

#define MAX_WORKERS 128

struct gpuhash_index_s {
	gpuhash_geom_t geom;
	void *table;
	int workers;
	size_t max_search, max_insert, max_delete;
	cudaStream_t stream[MAX_WORKERS];
	void *search_in_d[MAX_WORKERS];
	void *search_out_d[MAX_WORKERS];
	void *delete_in_d[MAX_WORKERS];
	void *insert_in_d[MAX_WORKERS];
	gpuhash_stats_t *stats_d;
	int stats_on;
	int zero_copy;       /* kernels read requests from / write results to the caller's pinned host buffers directly */
	int compact;         /* search_out_h gets ONE word per request (the sender's choice, mega_send.c:411-414) instead of two */
};

extern "C" void gpuhash_index_destroy(gpuhash_index_t *ix)
{
	if (!ix) return;
	cudaDeviceSynchronize();
	for (int w = 0; w < ix->workers; w++) {
		if (ix->stream[w]) cudaStreamDestroy(ix->stream[w]);
		cudaFree(ix->search_in_d[w]); cudaFree(ix->search_out_d[w]);
		cudaFree(ix->delete_in_d[w]); cudaFree(ix->insert_in_d[w]);
	}
	cudaFree(ix->stats_d);
	cudaFree(ix->table);
	free(ix);
}

extern "C" gpuhash_index_t *gpuhash_index_create_layout(int mem_p, unsigned algo, unsigned layout, int workers,
		size_t max_search, size_t max_insert, size_t max_delete);

extern "C" gpuhash_index_t *gpuhash_index_create(int mem_p, unsigned algo, int workers,
		size_t max_search, size_t max_insert, size_t max_delete)
{
	return gpuhash_index_create_layout(mem_p, algo, GPUHASH_LAYOUT_PAIRS, workers, max_search, max_insert, max_delete);
}

extern "C" gpuhash_index_t *gpuhash_index_create_layout(int mem_p, unsigned algo, unsigned layout, int workers,
		size_t max_search, size_t max_insert, size_t max_delete)
{
	if (workers < 1 || workers > MAX_WORKERS || layout > GPUHASH_LAYOUT_REFERENCE) return NULL;
	if (gpuhash_init_device() != 0) return NULL;
	gpuhash_index_t *ix = (gpuhash_index_t *)calloc(1, sizeof *ix);
	if (!ix) return NULL;
	if (gpuhash_geom_init(&ix->geom, mem_p, algo) != 0) { free(ix); return NULL; }
	ix->geom.layout = layout;
	ix->workers = workers;
	ix->max_search = max_search; ix->max_insert = max_insert; ix->max_delete = max_delete;
	size_t bytes = gpuhash_table_bytes(&ix->geom);
	int ok = cudaMalloc(&ix->table, bytes) == cudaSuccess
	      && cudaMemset(ix->table, 0, bytes) == cudaSuccess            /* all-zero == empty, mega_scheduler.c:273-274 */
	      && cudaMalloc((void **)&ix->stats_d, sizeof(gpuhash_stats_t)) == cudaSuccess
	      && cudaMemset(ix->stats_d, 0, sizeof(gpuhash_stats_t)) == cudaSuccess;
	for (int w = 0; ok && w < workers; w++) {
		ok = cudaStreamCreateWithFlags(&ix->stream[w], cudaStreamNonBlocking) == cudaSuccess
		  && cudaMalloc(&ix->search_in_d[w], (max_search ? max_search : 1) * 8) == cudaSuccess
		  && cudaMalloc(&ix->search_out_d[w], (max_search ? max_search : 1) * 8) == cudaSuccess
		  && cudaMalloc(&ix->delete_in_d[w], (max_delete ? max_delete : 1) * 12) == cudaSuccess
		  && cudaMalloc(&ix->insert_in_d[w], (max_insert ? max_insert : 1) * 12) == cudaSuccess;
	}
	if (!ok) {
		fprintf(stderr, "gpuhash_index_create: %s\n", cudaGetErrorString(cudaGetLastError()));
		gpuhash_index_destroy(ix);
		return NULL;
	}
	return ix;
}

extern "C" void *gpuhash_index_table(gpuhash_index_t *ix) { return ix->table; }
extern "C" const gpuhash_geom_t *gpuhash_index_geom(const gpuhash_index_t *ix) { return &ix->geom; }
extern "C" void *gpuhash_index_stream(gpuhash_index_t *ix, int w) { return (w >= 0 && w < ix->workers) ? (void *)ix->stream[w] : NULL; }

extern "C" int gpuhash_index_clear(gpuhash_index_t *ix)
{
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) return (int)e;
	return (int)cudaMemset(ix->table, 0, gpuhash_table_bytes(&ix->geom));
}

/* Host images are always in the reference's byte layout (bucket_t[], gpu_hash.h:79-82); the device table
 * is converted in place on the way in and out when it uses the pair layout. */
extern "C" int gpuhash_index_load(gpuhash_index_t *ix, const void *table_h)
{
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) return (int)e;
	e = cudaMemcpy(ix->table, table_h, gpuhash_table_bytes(&ix->geom), cudaMemcpyHostToDevice);
	if (e != cudaSuccess) return (int)e;
	if (ix->geom.layout != GPUHASH_LAYOUT_REFERENCE) {
		gpuhash_geom_t as_ref = ix->geom; as_ref.layout = GPUHASH_LAYOUT_REFERENCE;
		int rc = gpuhash_table_convert(&as_ref, ix->table, ix->geom.layout, NULL);
		if (rc) return rc;
	}
	return (int)cudaDeviceSynchronize();
}

extern "C" int gpuhash_index_dump(gpuhash_index_t *ix, void *table_h)
{
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) return (int)e;
	int rc = gpuhash_table_convert(&ix->geom, ix->table, GPUHASH_LAYOUT_REFERENCE, NULL);
	if (rc) return rc;
	e = cudaMemcpy(table_h, ix->table, gpuhash_table_bytes(&ix->geom), cudaMemcpyDeviceToHost);
	if (ix->geom.layout != GPUHASH_LAYOUT_REFERENCE) {
		gpuhash_geom_t as_ref = ix->geom; as_ref.layout = GPUHASH_LAYOUT_REFERENCE;
		rc = gpuhash_table_convert(&as_ref, ix->table, ix->geom.layout, NULL);
		if (rc) return rc;
	}
	if (e != cudaSuccess) return (int)e;
	return (int)cudaDeviceSynchronize();
}

extern "C" int gpuhash_index_enable_stats(gpuhash_index_t *ix, int on) { ix->stats_on = on != 0; return 0; }

/* Zero-copy mode: the host buffers handed to gpuhash_index_submit must be pinned (cudaHostAlloc / cudaHostRegister,
 * like the reference's batch buffers, mega_recv.c:154-156,176).  The kernels then read the requests and write the
 * results over PCIe themselves: no staging copy, no copy-engine call, one launch per non-empty part of the batch. */
extern "C" int gpuhash_index_set_zero_copy(gpuhash_index_t *ix, int on) { ix->zero_copy = on != 0; return 0; }
/* on: gpuhash_index_submit stores n_search words into search_out_h, word i = first non-zero of the reference's
 * {out[2i], out[2i+1]} -- the choice the sender makes (mega_send.c:411-414), made on the device: half the result bytes. */
extern "C" int gpuhash_index_set_compact_results(gpuhash_index_t *ix, int on) { ix->compact = on != 0; return 0; }

extern "C" int gpuhash_index_stats(gpuhash_index_t *ix, gpuhash_stats_t *out, int reset)
{
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) return (int)e;
	if (out && (e = cudaMemcpy(out, ix->stats_d, sizeof *out, cudaMemcpyDeviceToHost)) != cudaSuccess) return (int)e;
	if (reset) e = cudaMemset(ix->stats_d, 0, sizeof(gpuhash_stats_t));
	return (int)e;
}

extern "C" int gpuhash_index_submit(gpuhash_index_t *ix, int w,
		const void *search_in_h, size_t n_search, void *search_out_h,
		const void *delete_in_h, size_t n_delete,
		const void *insert_in_h, size_t n_insert)
{
	if (!ix || w < 0 || w >= ix->workers) return -1;
	if (n_search > ix->max_search || n_delete > ix->max_delete || n_insert > ix->max_insert) return -1;
	cudaStream_t s = ix->stream[w];
	gpuhash_stats_t *st = ix->stats_on ? ix->stats_d : NULL;
	cudaError_t e;
	int rc;
	gpuhash_tune_t tune; gpuhash_get_tuning(&tune);
	if (ix->compact) {                                                         /* one result word per search; one launch per operation kind */
		if (n_search) {
			if (ix->zero_copy) {
				if ((rc = gpuhash_search_compact_ex(&ix->geom, search_in_h, search_out_h, ix->table, n_search, st, s)) != 0) return rc;
			} else {
				if ((e = cudaMemcpyAsync(ix->search_in_d[w], search_in_h, n_search * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
				if ((rc = gpuhash_search_compact_ex(&ix->geom, ix->search_in_d[w], ix->search_out_d[w], ix->table, n_search, st, s)) != 0) return rc;
				if ((e = cudaMemcpyAsync(search_out_h, ix->search_out_d[w], n_search * 4, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
			}
		}
		if (n_delete) {
			const void *din = delete_in_h;
			if (!ix->zero_copy) { if ((e = cudaMemcpyAsync(ix->delete_in_d[w], delete_in_h, n_delete * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e; din = ix->delete_in_d[w]; }
			if ((rc = gpuhash_delete_ex(&ix->geom, din, ix->table, n_delete, st, 0, s)) != 0) return rc;
		}
		if (n_insert) {
			const void *iin = insert_in_h;
			if (!ix->zero_copy) { if ((e = cudaMemcpyAsync(ix->insert_in_d[w], insert_in_h, n_insert * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e; iin = ix->insert_in_d[w]; }
			if ((rc = gpuhash_insert_flat_ex(&ix->geom, ix->table, iin, n_insert, st, 0, s)) != 0) return rc;
		}
		return 0;
	}
	if (ix->zero_copy) {
		if (tune.fused_cycle)
			return gpuhash_cycle_ex(&ix->geom, ix->table, search_in_h, n_search, search_out_h, delete_in_h, n_delete,
					insert_in_h, n_insert, NULL, NULL, 0, st, s);
		if (n_search && (rc = gpuhash_search_ex(&ix->geom, search_in_h, search_out_h, ix->table, n_search, st, s)) != 0) return rc;
		if (n_delete && (rc = gpuhash_delete_ex(&ix->geom, delete_in_h, ix->table, n_delete, st, 0, s)) != 0) return rc;
		if (n_insert && (rc = gpuhash_insert_flat_ex(&ix->geom, ix->table, insert_in_h, n_insert, st, 0, s)) != 0) return rc;
		return 0;
	}
	if (tune.fused_cycle) {                                                    /* all uploads, ONE launch, the download */
		if (n_search && (e = cudaMemcpyAsync(ix->search_in_d[w], search_in_h, n_search * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if (n_delete && (e = cudaMemcpyAsync(ix->delete_in_d[w], delete_in_h, n_delete * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if (n_insert && (e = cudaMemcpyAsync(ix->insert_in_d[w], insert_in_h, n_insert * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if ((rc = gpuhash_cycle_ex(&ix->geom, ix->table, ix->search_in_d[w], n_search, ix->search_out_d[w], ix->delete_in_d[w], n_delete,
				ix->insert_in_d[w], n_insert, NULL, NULL, 0, st, s)) != 0) return rc;
		if (n_search && (e = cudaMemcpyAsync(search_out_h, ix->search_out_d[w], n_search * 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
		return 0;
	}
	if (n_search) {                                                            /* mega_scheduler.c:393-420 */
		if ((e = cudaMemcpyAsync(ix->search_in_d[w], search_in_h, n_search * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if ((rc = gpuhash_search_ex(&ix->geom, ix->search_in_d[w], ix->search_out_d[w], ix->table, n_search, st, s)) != 0) return rc;
		if ((e = cudaMemcpyAsync(search_out_h, ix->search_out_d[w], n_search * 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return (int)e;
	}
	if (n_delete) {                                                            /* mega_scheduler.c:440-462 */
		if ((e = cudaMemcpyAsync(ix->delete_in_d[w], delete_in_h, n_delete * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if ((rc = gpuhash_delete_ex(&ix->geom, ix->delete_in_d[w], ix->table, n_delete, st, 0, s)) != 0) return rc;
	}
	if (n_insert) {                                                            /* mega_scheduler.c:472-502 */
		if ((e = cudaMemcpyAsync(ix->insert_in_d[w], insert_in_h, n_insert * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if ((rc = gpuhash_insert_flat_ex(&ix->geom, ix->table, ix->insert_in_d[w], n_insert, st, 0, s)) != 0) return rc;
	}
	return 0;
}

extern "C" int gpuhash_index_sync(gpuhash_index_t *ix)
{
	(void)ix;
	return (int)cudaDeviceSynchronize();                                       /* mega_scheduler.c:504 */
}

/* ------------------------------------------------------------------ timed loops */

static int issue_resident(const gpuhash_geom_t *g, void *table_d,
		const char *search_d, size_t n_search, char *out_d, const char *insert_d, size_t n_insert,
		int steps, cudaStream_t *st, int streams)
{
	gpuhash_tune_t tune; gpuhash_get_tuning(&tune);
	for (int i = 0; i < steps; i++) {
		cudaStream_t s = st[i % streams];
		int rc;
		if (tune.fused_cycle && n_search && n_insert) {
			if ((rc = gpuhash_cycle_ex(g, table_d, search_d + (size_t)i * n_search * 8, n_search, out_d + (size_t)i * n_search * 8,
					NULL, 0, insert_d + (size_t)i * n_insert * 12, n_insert, NULL, NULL, 0, NULL, s)) != 0) return rc;
			continue;
		}
		if (n_search && (rc = gpuhash_search_ex(g, search_d + (size_t)i * n_search * 8, out_d + (size_t)i * n_search * 8,
				table_d, n_search, NULL, s)) != 0) return rc;
		if (n_insert && (rc = gpuhash_insert_flat_ex(g, table_d, insert_d + (size_t)i * n_insert * 12, n_insert,
				NULL, 0, s)) != 0) return rc;
	}
	return 0;
}

extern "C" int gpuhash_bench_resident(const gpuhash_geom_t *g, void *table_d,
		const void *search_d, size_t n_search, void *out_d,
		const void *insert_d, size_t n_insert,
		int steps, int streams, int use_graph, gpuhash_bench_result_t *res)
{
	if (!g || !res || steps < 1 || streams < 1 || streams > MAX_WORKERS) return -1;
	if (gpuhash_init_device() != 0) return -1;
	memset(res, 0, sizeof *res);
	cudaStream_t st[MAX_WORKERS], main_s;
	cudaEvent_t ev_start, ev_stop, ev_done[MAX_WORKERS];
	cudaError_t e = cudaSuccess;
	int rc = 0;
	cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking);
	for (int k = 0; k < streams; k++) { cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking); cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming); }
	cudaEventCreate(&ev_start); cudaEventCreate(&ev_stop);
	cudaGraph_t graph = NULL; cudaGraphExec_t gexec = NULL;

	if (use_graph) {
		/* fork: main_s -> st[k]; issue; join: st[k] -> main_s; all inside one capture */
		cudaEvent_t fork; cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
		e = cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal);
		if (e == cudaSuccess) {
			cudaEventRecord(fork, main_s);
			for (int k = 0; k < streams; k++) cudaStreamWaitEvent(st[k], fork, 0);
			rc = issue_resident(g, table_d, (const char *)search_d, n_search, (char *)out_d, (const char *)insert_d, n_insert, steps, st, streams);
			for (int k = 0; k < streams; k++) { cudaEventRecord(ev_done[k], st[k]); cudaStreamWaitEvent(main_s, ev_done[k], 0); }
			e = cudaStreamEndCapture(main_s, &graph);
		}
		if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&gexec, graph, 0);
		cudaEventDestroy(fork);
		if (e == cudaSuccess && rc == 0) {
			cudaDeviceSynchronize();
			cudaEventRecord(ev_start, main_s);
			e = cudaGraphLaunch(gexec, main_s);
			cudaEventRecord(ev_stop, main_s);
		}
	} else {
		cudaDeviceSynchronize();
		cudaEventRecord(ev_start, main_s);
		for (int k = 0; k < streams; k++) cudaStreamWaitEvent(st[k], ev_start, 0);
		rc = issue_resident(g, table_d, (const char *)search_d, n_search, (char *)out_d, (const char *)insert_d, n_insert, steps, st, streams);
		for (int k = 0; k < streams; k++) { cudaEventRecord(ev_done[k], st[k]); cudaStreamWaitEvent(main_s, ev_done[k], 0); }
		cudaEventRecord(ev_stop, main_s);
	}
	if (e == cudaSuccess && rc == 0) e = cudaEventSynchronize(ev_stop);
	if (e == cudaSuccess && rc == 0) cudaEventElapsedTime(&res->total_ms, ev_start, ev_stop);
	cudaDeviceSynchronize();
	if (gexec) cudaGraphExecDestroy(gexec);
	if (graph) cudaGraphDestroy(graph);
	for (int k = 0; k < streams; k++) { cudaStreamDestroy(st[k]); cudaEventDestroy(ev_done[k]); }
	cudaStreamDestroy(main_s); cudaEventDestroy(ev_start); cudaEventDestroy(ev_stop);
	res->launches = (unsigned long long)steps * ((n_search ? 1 : 0) + (n_insert ? 1 : 0));
	res->search_ops = (unsigned long long)steps * n_search;
	res->insert_ops = (unsigned long long)steps * n_insert;
	if (rc != 0) return rc;
	return (int)e;
}

/* K cycles through the host-buffer API.  Cycle i uses worker slot i % workers and the i-th batch of the
 * pinned arrays (search_h: steps*n_search selem_t, out_h: steps*2*n_search loc_t, insert_h: steps*n_insert
 * ielem_t), so every step moves its own bytes over PCIe in both directions inside the timed region.
 * use_graph: the submits are captured once into a CUDA graph (copies and kernels become graph nodes) and the
 * graph launch is timed -- what a scheduler with fixed pinned batch buffers would replay every cycle. */
extern "C" int gpuhash_bench_e2e(gpuhash_index_t *ix,
		const void *search_h, size_t n_search, void *out_h,
		const void *insert_h, size_t n_insert,
		int steps, int use_graph, gpuhash_bench_result_t *res)
{
	if (!ix || !res || steps < 1) return -1;
	memset(res, 0, sizeof *res);
	cudaStream_t main_s;
	cudaEvent_t ev_start, ev_stop, ev_done[MAX_WORKERS], fork;
	cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking);
	cudaEventCreate(&ev_start); cudaEventCreate(&ev_stop); cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
	for (int k = 0; k < ix->workers; k++) cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming);
	cudaGraph_t graph = NULL; cudaGraphExec_t gexec = NULL;
	cudaError_t e = cudaSuccess;
	int rc = 0;
	cudaDeviceSynchronize();
	if (use_graph) e = cudaStreamBeginCapture(main_s, cudaStreamCaptureModeThreadLocal);
	else cudaEventRecord(ev_start, main_s);
	if (e == cudaSuccess) {
		cudaEventRecord(fork, main_s);
		for (int k = 0; k < ix->workers; k++) cudaStreamWaitEvent(ix->stream[k], fork, 0);
		for (int i = 0; i < steps && rc == 0; i++)
			rc = gpuhash_index_submit(ix, i % ix->workers,
					(const char *)search_h + (size_t)i * n_search * 8, n_search, (char *)out_h + (size_t)i * n_search * (ix->compact ? 4 : 8),
					NULL, 0,
					(const char *)insert_h + (size_t)i * n_insert * 12, n_insert);
		for (int k = 0; k < ix->workers; k++) { cudaEventRecord(ev_done[k], ix->stream[k]); cudaStreamWaitEvent(main_s, ev_done[k], 0); }
		if (use_graph) {
			e = cudaStreamEndCapture(main_s, &graph);
			if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&gexec, graph, 0);
			if (e == cudaSuccess && rc == 0) {
				cudaDeviceSynchronize();
				cudaEventRecord(ev_start, main_s);
				e = cudaGraphLaunch(gexec, main_s);
			}
		}
		cudaEventRecord(ev_stop, main_s);
	}
	if (e == cudaSuccess && rc == 0) e = cudaEventSynchronize(ev_stop);
	if (e == cudaSuccess && rc == 0) cudaEventElapsedTime(&res->total_ms, ev_start, ev_stop);
	cudaDeviceSynchronize();
	if (gexec) cudaGraphExecDestroy(gexec);
	if (graph) cudaGraphDestroy(graph);
	for (int k = 0; k < ix->workers; k++) cudaEventDestroy(ev_done[k]);
	cudaStreamDestroy(main_s); cudaEventDestroy(ev_start); cudaEventDestroy(ev_stop); cudaEventDestroy(fork);
	res->launches = (unsigned long long)steps * ((n_search ? 1 : 0) + (n_insert ? 1 : 0));
	res->search_ops = (unsigned long long)steps * n_search;
	res->insert_ops = (unsigned long long)steps * n_insert;
	res->h2d_bytes = (unsigned long long)steps * (n_search * 8 + n_insert * 12);
	res->d2h_bytes = (unsigned long long)steps * n_search * (ix->compact ? 4 : 8);
	if (rc != 0) return rc;
	return (int)e;
}
