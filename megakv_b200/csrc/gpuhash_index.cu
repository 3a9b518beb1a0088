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

struct cycle_slot {                 /* one whole-cycle launch in flight (gpuhash_index_submit_all) */
	cudaStream_t stream;
	cudaEvent_t done;
	gpuhash_batch_t *desc_h;        /* pinned mirror of the descriptor table the kernel reads */
	gpuhash_batch_t *desc_d;
	void *ws_d;                     /* gpuhash_cycle_workspace_bytes(MAX_WORKERS), zero between launches */
	/* staged mode: this slot's own device copies of all workers' batches, packed back to back in worker order (allocated at the
	 * first staged submit): cycles of different slots overlap -- H2D of one, kernel of the next, D2H of a third */
	char *st_search, *st_out, *st_delete, *st_insert;
	cudaEvent_t kernel_done;        /* recorded right behind the cycle kernel: the next cycle's kernel waits for it (ordered mode) */
	int launched;
	int busy;
};

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
	void *ws_d[MAX_WORKERS];        /* one-launch cycle of worker w: its own workspace (launches of a stream are serial) */
	struct cycle_slot slot[GPUHASH_INDEX_SLOTS];
	unsigned next_slot;
	gpuhash_stats_t *stats_d;
	int stats_on;
	int zero_copy;       /* kernels read requests from / write results to the caller's pinned host buffers directly */
	int compact;         /* search_out_h gets ONE word per request (the sender's choice, mega_send.c:411-414) instead of two */
	int unordered;       /* 0 (default): the kernel of cycle k+1 starts after the kernel of cycle k has finished, as the reference's
	                        cycles do (one cudaDeviceSynchronize each, mega_scheduler.c:504): a GET in cycle k+1 sees a SET of cycle k */
};

extern "C" void gpuhash_index_destroy(gpuhash_index_t *ix)
{
	if (!ix) return;
	cudaDeviceSynchronize();
	for (int w = 0; w < ix->workers; w++) {
		if (ix->stream[w]) cudaStreamDestroy(ix->stream[w]);
		cudaFree(ix->search_in_d[w]); cudaFree(ix->search_out_d[w]);
		cudaFree(ix->delete_in_d[w]); cudaFree(ix->insert_in_d[w]); cudaFree(ix->ws_d[w]);
	}
	for (int k = 0; k < GPUHASH_INDEX_SLOTS; k++) {
		struct cycle_slot *c = &ix->slot[k];
		if (c->stream) cudaStreamDestroy(c->stream);
		if (c->done) cudaEventDestroy(c->done);
		if (c->kernel_done) cudaEventDestroy(c->kernel_done);
		if (c->desc_h) cudaFreeHost(c->desc_h);
		cudaFree(c->desc_d); cudaFree(c->ws_d);
		cudaFree(c->st_search); cudaFree(c->st_out); cudaFree(c->st_delete); cudaFree(c->st_insert);
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
	const size_t ws1 = gpuhash_cycle_workspace_bytes(1), wsn = gpuhash_cycle_workspace_bytes(MAX_WORKERS);
	for (int w = 0; ok && w < workers; w++) {
		ok = cudaStreamCreateWithFlags(&ix->stream[w], cudaStreamNonBlocking) == cudaSuccess
		  && cudaMalloc(&ix->search_in_d[w], (max_search ? max_search : 1) * 8) == cudaSuccess
		  && cudaMalloc(&ix->search_out_d[w], (max_search ? max_search : 1) * 8) == cudaSuccess
		  && cudaMalloc(&ix->delete_in_d[w], (max_delete ? max_delete : 1) * 12) == cudaSuccess
		  && cudaMalloc(&ix->insert_in_d[w], (max_insert ? max_insert : 1) * 12) == cudaSuccess
		  && cudaMalloc(&ix->ws_d[w], ws1) == cudaSuccess
		  && cudaMemset(ix->ws_d[w], 0, ws1) == cudaSuccess;
	}
	for (int k = 0; ok && k < GPUHASH_INDEX_SLOTS; k++) {
		struct cycle_slot *c = &ix->slot[k];
		ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess
		  && cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming) == cudaSuccess
		  && cudaEventCreateWithFlags(&c->kernel_done, cudaEventDisableTiming) == cudaSuccess
		  && cudaHostAlloc((void **)&c->desc_h, sizeof(gpuhash_batch_t) * MAX_WORKERS, cudaHostAllocDefault) == cudaSuccess
		  && cudaMalloc((void **)&c->desc_d, sizeof(gpuhash_batch_t) * MAX_WORKERS) == cudaSuccess
		  && cudaMalloc(&c->ws_d, wsn) == cudaSuccess
		  && cudaMemset(c->ws_d, 0, wsn) == cudaSuccess;
	}
	/* the fills above are asynchronous to the host and every stream of the index is non-blocking (no implicit order
	 * against the legacy stream they ran on): nothing may be submitted before they have finished */
	if (ok) ok = cudaDeviceSynchronize() == cudaSuccess;
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
	if ((e = cudaMemset(ix->table, 0, gpuhash_table_bytes(&ix->geom))) != cudaSuccess) return (int)e;
	return (int)cudaDeviceSynchronize();          /* the fill is asynchronous to the host; the index's streams are non-blocking */
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
	if (reset && (e = cudaMemset(ix->stats_d, 0, sizeof(gpuhash_stats_t))) == cudaSuccess) e = cudaDeviceSynchronize();
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
			return gpuhash_cycle_ws_ex(&ix->geom, ix->table, search_in_h, n_search, search_out_h, delete_in_h, n_delete,
					insert_in_h, n_insert, NULL, NULL, 0, 0, ix->ws_d[w], st, s);
		if (n_search && (rc = gpuhash_search_ex(&ix->geom, search_in_h, search_out_h, ix->table, n_search, st, s)) != 0) return rc;
		if (n_delete && (rc = gpuhash_delete_ex(&ix->geom, delete_in_h, ix->table, n_delete, st, 0, s)) != 0) return rc;
		if (n_insert && (rc = gpuhash_insert_flat_ex(&ix->geom, ix->table, insert_in_h, n_insert, st, 0, s)) != 0) return rc;
		return 0;
	}
	if (tune.fused_cycle) {                                                    /* all uploads, ONE launch, the download */
		if (n_search && (e = cudaMemcpyAsync(ix->search_in_d[w], search_in_h, n_search * 8, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if (n_delete && (e = cudaMemcpyAsync(ix->delete_in_d[w], delete_in_h, n_delete * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if (n_insert && (e = cudaMemcpyAsync(ix->insert_in_d[w], insert_in_h, n_insert * 12, cudaMemcpyHostToDevice, s)) != cudaSuccess) return (int)e;
		if ((rc = gpuhash_cycle_ws_ex(&ix->geom, ix->table, ix->search_in_d[w], n_search, ix->search_out_d[w], ix->delete_in_d[w], n_delete,
				ix->insert_in_d[w], n_insert, NULL, NULL, 0, 0, ix->ws_d[w], st, s)) != 0) return rc;
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

static int cycle_error(gpuhash_index_t *ix)
{
	(void)ix;
	return gpuhash_cycle_error(1) ? -3 : 0;
}

extern "C" int gpuhash_index_sync(gpuhash_index_t *ix)
{
	cudaError_t e = cudaDeviceSynchronize();                                   /* mega_scheduler.c:504 */
	for (int k = 0; k < GPUHASH_INDEX_SLOTS; k++) ix->slot[k].busy = 0;
	if (e != cudaSuccess) return (int)e;
	return cycle_error(ix);
}

/* mega_scheduler.c:393-502 for all workers at once: one descriptor upload, ONE launch (zero-copy), one event.
 * Staged mode (pageable or pinned host buffers): every slot owns device copies of all workers' batches, packed back to back
 * in worker order, so the cycles in flight overlap (copy-in of one, kernel of the next, copy-out of a third), and host
 * buffers that are ADJACENT in memory -- worker w+1's array starting where worker w's ends, as when the receiver carves all
 * its batch buffers out of one pinned block -- travel as ONE copy: the copy engines reach 48-55 GB/s per direction on
 * multi-megabyte copies against 24 GB/s on the 0.5 MB copies of single batches (tools/pcie_probe). */
static int slot_stage_alloc(gpuhash_index_t *ix, struct cycle_slot *c)
{
	if (c->st_search) return 0;
	const size_t W = (size_t)ix->workers;
	if (cudaMalloc((void **)&c->st_search, W * (ix->max_search ? ix->max_search : 1) * 8) != cudaSuccess
	 || cudaMalloc((void **)&c->st_out, W * (ix->max_search ? ix->max_search : 1) * 8) != cudaSuccess
	 || cudaMalloc((void **)&c->st_delete, W * (ix->max_delete ? ix->max_delete : 1) * 12) != cudaSuccess
	 || cudaMalloc((void **)&c->st_insert, W * (ix->max_insert ? ix->max_insert : 1) * 12) != cudaSuccess) return -1;
	return 0;
}

/* copy `num` pieces (host h[w], bytes n[w], device d[w] = packed) merging neighbours that are adjacent on the host */
static cudaError_t copy_runs(int to_device, char *const *dev, char *const *host, const size_t *bytes, int num, cudaStream_t s)
{
	int w = 0;
	while (w < num) {
		if (bytes[w] == 0) { w++; continue; }
		size_t run = bytes[w];
		int v = w + 1;
		while (v < num && (bytes[v] == 0 || (host[v] == host[w] + run && dev[v] == dev[w] + run))) { run += bytes[v]; v++; }
		cudaError_t e = to_device ? cudaMemcpyAsync(dev[w], host[w], run, cudaMemcpyHostToDevice, s)
		                          : cudaMemcpyAsync(host[w], dev[w], run, cudaMemcpyDeviceToHost, s);
		if (e != cudaSuccess) return e;
		w = v;
	}
	return cudaSuccess;
}

extern "C" int gpuhash_index_submit_all(gpuhash_index_t *ix, const gpuhash_batch_t *batches_h, int num_batches)
{
	if (!ix || !batches_h || num_batches < 1 || num_batches > ix->workers) return -1;
	for (int w = 0; w < num_batches; w++)
		if (batches_h[w].n_search > ix->max_search || batches_h[w].n_delete > ix->max_delete || batches_h[w].n_insert > ix->max_insert) return -1;
	const unsigned k = ix->next_slot % GPUHASH_INDEX_SLOTS;
	struct cycle_slot *c = &ix->slot[k];
	cudaError_t e;
	if (c->busy && (e = cudaEventSynchronize(c->done)) != cudaSuccess) return -(int)e - 16;   /* descriptor mirror and workspace still in use */
	c->busy = 0;
	cudaStream_t s = c->stream;
	gpuhash_stats_t *st = ix->stats_on ? ix->stats_d : NULL;
	const size_t out_bytes = ix->compact ? 4 : 8;
	char *dv[4][MAX_WORKERS], *hv[4][MAX_WORKERS]; size_t nb[4][MAX_WORKERS];      /* search in, delete in, insert in, search out */
	if (ix->zero_copy) {
		memcpy(c->desc_h, batches_h, sizeof(gpuhash_batch_t) * (size_t)num_batches);
	} else {
		if (slot_stage_alloc(ix, c)) return -(int)cudaErrorMemoryAllocation - 16;
		size_t os = 0, od = 0, oi = 0;
		for (int w = 0; w < num_batches; w++) {
			const gpuhash_batch_t *b = &batches_h[w];
			gpuhash_batch_t *d = &c->desc_h[w];
			memset(d, 0, sizeof *d);
			d->n_search = b->n_search; d->n_delete = b->n_delete; d->n_insert = b->n_insert;
			dv[0][w] = c->st_search + os; dv[1][w] = c->st_delete + od; dv[2][w] = c->st_insert + oi;
			dv[3][w] = c->st_out + os / 8 * out_bytes;
			hv[0][w] = (char *)b->search_in; hv[1][w] = (char *)b->delete_in; hv[2][w] = (char *)b->insert_in; hv[3][w] = (char *)b->search_out;
			nb[0][w] = (size_t)b->n_search * 8; nb[1][w] = (size_t)b->n_delete * 12; nb[2][w] = (size_t)b->n_insert * 12;
			nb[3][w] = (size_t)b->n_search * out_bytes;
			d->search_in = dv[0][w]; d->search_out = dv[3][w]; d->delete_in = dv[1][w]; d->insert_in = dv[2][w];
			os += nb[0][w]; od += nb[1][w]; oi += nb[2][w];
		}
		for (int kind = 0; kind < 3; kind++)
			if ((e = copy_runs(1, dv[kind], hv[kind], nb[kind], num_batches, s)) != cudaSuccess) return -(int)e - 16;
	}
	if ((e = cudaMemcpyAsync(c->desc_d, c->desc_h, sizeof(gpuhash_batch_t) * (size_t)num_batches, cudaMemcpyHostToDevice, s)) != cudaSuccess) return -(int)e - 16;
	if (!ix->unordered) {                                /* cycles keep their order: this kernel behind the previous cycle's kernel */
		const struct cycle_slot *prev = &ix->slot[(k + GPUHASH_INDEX_SLOTS - 1) % GPUHASH_INDEX_SLOTS];
		if (prev->launched && (e = cudaStreamWaitEvent(s, prev->kernel_done, 0)) != cudaSuccess) return -(int)e - 16;
	}
	int rc = gpuhash_cycle_multi_ex(&ix->geom, ix->table, c->desc_h, c->desc_d, num_batches, ix->compact, c->ws_d, st, s);
	if (rc) return rc < 0 ? rc : -rc - 16;
	if ((e = cudaEventRecord(c->kernel_done, s)) != cudaSuccess) return -(int)e - 16;
	c->launched = 1;
	if (!ix->zero_copy && (e = copy_runs(0, dv[3], hv[3], nb[3], num_batches, s)) != cudaSuccess) return -(int)e - 16;
	if ((e = cudaEventRecord(c->done, s)) != cudaSuccess) return -(int)e - 16;
	c->busy = 1;
	ix->next_slot++;
	return (int)k;
}

extern "C" int gpuhash_index_set_unordered_cycles(gpuhash_index_t *ix, int on) { if (!ix) return -1; ix->unordered = on != 0; return 0; }

extern "C" int gpuhash_index_wait(gpuhash_index_t *ix, int ticket)
{
	if (!ix || ticket < 0 || ticket >= GPUHASH_INDEX_SLOTS) return -1;
	struct cycle_slot *c = &ix->slot[ticket];
	if (c->busy) {
		cudaError_t e = cudaEventSynchronize(c->done);
		c->busy = 0;
		if (e != cudaSuccess) return (int)e;
	}
	return cycle_error(ix);
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
					NULL, 0, insert_d + (size_t)i * n_insert * 12, n_insert, NULL, NULL, 0, NULL, s)) != 0) return rc;   /* pool workspace: direct calls or ONE replayed graph at a time */
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

/* ------------------------------------------------------------------ whole cycles: one launch per step */

#include <time.h>
static double wall_ms_now(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

/* `steps` scheduler cycles, each ONE launch (gpuhash_cycle_multi_ex) over `batches` worker batches resident in device
 * memory: batch b of step i is batch number i * batches + b of search_d / out_d / insert_d (n_search selem_t, 2 n_search
 * loc_t, n_insert ielem_t each).  Steps go round-robin over `streams` (1..4) streams, each with its own workspace: with
 * one stream the cycles run strictly one after the other, like the reference's (one cudaDeviceSynchronize per cycle,
 * mega_scheduler.c:504); with two the tail of a cycle overlaps the head of the next, which its triple-buffered batches
 * allow (mega_batch.h:74-82).  Descriptor tables are uploaded before the timed region; CUDA events around the launches. */
extern "C" int gpuhash_bench_cycles(const gpuhash_geom_t *g, void *table_d,
		const void *search_d, size_t n_search, void *out_d,
		const void *insert_d, size_t n_insert,
		int batches, int steps, int streams, gpuhash_bench_result_t *res)
{
	if (!g || !res || steps < 1 || batches < 1 || batches > GPUHASH_MAX_BATCHES || streams < 1 || streams > 4) return -1;
	if (gpuhash_init_device() != 0) return -1;
	memset(res, 0, sizeof *res);
	const size_t nd = (size_t)steps * (size_t)batches;
	gpuhash_batch_t *dh = (gpuhash_batch_t *)calloc(nd, sizeof *dh), *dd = NULL;
	if (!dh) return -1;
	for (size_t k = 0; k < nd; k++) {
		dh[k].search_in = n_search ? (const char *)search_d + k * n_search * 8 : NULL;
		dh[k].search_out = n_search ? (char *)out_d + k * n_search * 8 : NULL;
		dh[k].insert_in = n_insert ? (const char *)insert_d + k * n_insert * 12 : NULL;
		dh[k].n_search = (uint32_t)n_search; dh[k].n_insert = (uint32_t)n_insert;
	}
	cudaStream_t st[4] = {0}, main_s = NULL;
	cudaEvent_t ev_start = NULL, ev_stop = NULL, ev_done[4] = {0};
	void *ws[4] = {0};
	const size_t wsb = gpuhash_cycle_workspace_bytes(batches);
	cudaError_t e = cudaMalloc((void **)&dd, nd * sizeof *dd);
	if (e == cudaSuccess) e = cudaMemcpy(dd, dh, nd * sizeof *dd, cudaMemcpyHostToDevice);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking);
	for (int k = 0; e == cudaSuccess && k < streams; k++) {
		e = cudaStreamCreateWithFlags(&st[k], cudaStreamNonBlocking);
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_done[k], cudaEventDisableTiming);
		if (e == cudaSuccess) e = cudaMalloc(&ws[k], wsb);
		if (e == cudaSuccess) e = cudaMemset(ws[k], 0, wsb);
	}
	if (e == cudaSuccess) e = cudaEventCreate(&ev_start);
	if (e == cudaSuccess) e = cudaEventCreate(&ev_stop);
	int rc = 0;
	if (e == cudaSuccess) {
		cudaDeviceSynchronize();
		cudaEventRecord(ev_start, main_s);
		for (int k = 0; k < streams; k++) cudaStreamWaitEvent(st[k], ev_start, 0);
		for (int i = 0; i < steps && rc == 0; i++)
			rc = gpuhash_cycle_multi_ex(g, table_d, dh + (size_t)i * batches, dd + (size_t)i * batches, batches, 0, ws[i % streams], NULL, st[i % streams]);
		for (int k = 0; k < streams; k++) { cudaEventRecord(ev_done[k], st[k]); cudaStreamWaitEvent(main_s, ev_done[k], 0); }
		cudaEventRecord(ev_stop, main_s);
		e = cudaEventSynchronize(ev_stop);
		if (e == cudaSuccess && rc == 0) cudaEventElapsedTime(&res->total_ms, ev_start, ev_stop);
	}
	cudaDeviceSynchronize();
	if (rc == 0 && gpuhash_cycle_error(1)) rc = -3;
	for (int k = 0; k < streams; k++) { if (st[k]) cudaStreamDestroy(st[k]); if (ev_done[k]) cudaEventDestroy(ev_done[k]); cudaFree(ws[k]); }
	if (main_s) cudaStreamDestroy(main_s);
	if (ev_start) cudaEventDestroy(ev_start);
	if (ev_stop) cudaEventDestroy(ev_stop);
	cudaFree(dd); free(dh);
	res->launches = (unsigned long long)steps;
	res->search_ops = (unsigned long long)nd * n_search;
	res->insert_ops = (unsigned long long)nd * n_insert;
	if (rc != 0) return rc;
	return (int)e;
}

/* The same cycles end to end through the host-buffer call a scheduler makes: gpuhash_index_submit_all per step (pinned
 * host batches in, results back in pinned host memory), at most `depth` (1..GPUHASH_INDEX_SLOTS) cycles in flight,
 * gpuhash_index_wait before a slot is reused and for everything at the end.  total_ms is the HOST'S WALL CLOCK from the
 * first submit to the return of the last wait: launch latency, descriptor uploads and synchronisation included.
 * Batch b of step i is batch (i * batches + b) % host_batches of the pinned arrays. */
extern "C" int gpuhash_bench_e2e_cycles(gpuhash_index_t *ix,
		const void *search_h, size_t n_search, void *out_h,
		const void *insert_h, size_t n_insert,
		int batches, size_t host_batches, int steps, int depth, gpuhash_bench_result_t *res)
{
	if (!ix || !res || steps < 1 || batches < 1 || batches > ix->workers || depth < 1 || depth > GPUHASH_INDEX_SLOTS) return -1;
	if (host_batches < (size_t)batches * (size_t)depth) return -1;
	memset(res, 0, sizeof *res);
	const size_t ob = ix->compact ? 4 : 8;
	gpuhash_batch_t *d = (gpuhash_batch_t *)calloc((size_t)batches, sizeof *d);
	if (!d) return -1;
	int tickets[GPUHASH_INDEX_SLOTS];
	int rc = 0;
	cudaDeviceSynchronize();
	const double t0 = wall_ms_now();
	for (int i = 0; i < steps && rc == 0; i++) {
		for (int b = 0; b < batches; b++) {
			const size_t k = ((size_t)i * batches + b) % host_batches;
			d[b].search_in = n_search ? (const char *)search_h + k * n_search * 8 : NULL;
			d[b].search_out = n_search ? (char *)out_h + k * n_search * ob : NULL;
			d[b].insert_in = n_insert ? (const char *)insert_h + k * n_insert * 12 : NULL;
			d[b].n_search = (uint32_t)n_search; d[b].n_insert = (uint32_t)n_insert;
		}
		if (i >= depth) rc = gpuhash_index_wait(ix, tickets[i % depth]);       /* the scheduler hands that batch on before reusing the slot */
		if (rc) break;
		const int t = gpuhash_index_submit_all(ix, d, batches);
		if (t < 0) { rc = t; break; }
		tickets[i % depth] = t;
	}
	for (int i = steps > depth ? steps - depth : 0; i < steps && rc == 0; i++) rc = gpuhash_index_wait(ix, tickets[i % depth]);
	const double t1 = wall_ms_now();
	cudaDeviceSynchronize();
	free(d);
	res->total_ms = (float)(t1 - t0);
	res->launches = (unsigned long long)steps;
	res->search_ops = (unsigned long long)steps * batches * n_search;
	res->insert_ops = (unsigned long long)steps * batches * n_insert;
	res->h2d_bytes = (unsigned long long)steps * batches * (n_search * 8 + n_insert * 12);
	res->d2h_bytes = (unsigned long long)steps * batches * n_search * ob;
	return rc;
}
