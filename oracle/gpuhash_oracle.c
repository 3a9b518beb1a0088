/*
 * gpuhash_oracle.c -- see gpuhash_oracle.h.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Each function names the reference lines (pzrq/megakv, libgpuhash/) it follows.
 * The reference lets 8 lanes look at the 8 slots of a bucket at once and picks
 * with __ffs(__ballot(...)); here a lane is a loop index and "lowest set bit"
 * is "first l for which the predicate holds".
 */
#define _GNU_SOURCE
#include "gpuhash_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SLOTS 8                       /* ELEM_NUM, gpu_hash.h:49 */

typedef struct bucket_s {             /* bucket_t, gpu_hash.h:79-82 */
	uint32_t sig[SLOTS];
	uint32_t loc[SLOTS];
} bucket_t;

void orc_geom_init(orc_geom_t *g, int mem_p, int algo)
{
	/* gpu_hash.h:57 BUC_P = 6; :61 HASH_MASK; :67-69 IBLOCK_P = 3, BLOCK_HASH_MASK */
	int buc_p = 6, iblock_p = 3;
	g->hash_mask  = (uint32_t)((1ULL << (mem_p - buc_p)) - 1);
	g->block_mask = (uint32_t)((1ULL << (mem_p - buc_p - iblock_p)) - 1);
	g->algo = algo;
	g->max_cuckoo = 5;                /* gpu_hash.h:75 */
}

size_t orc_table_bytes(const orc_geom_t *g)
{
	return ((size_t)g->hash_mask + 1) * sizeof(bucket_t);
}

/* gpu_hash.cu:55 */
uint32_t orc_bucket1(const orc_geom_t *g, uint32_t hash)
{
	return hash & g->hash_mask;
}

/* gpu_hash.cu:66-67 (search), :171-172, :334-335 (insert), :471-472 (delete) */
uint32_t orc_bucket2(const orc_geom_t *g, uint32_t hash, uint32_t sig)
{
	return (((hash ^ sig) & g->block_mask) | (hash & ~g->block_mask)) & g->hash_mask;
}

/* ------------------------------------------------------------------ search */

/* gpu_hash.cu:47-72.  Every lane whose slot matches stores its loc to the same
 * word; with more than one matching lane the hardware keeps one of the stores.
 * A table built by sequential inserts never holds a non-zero signature twice
 * in a bucket, so this only matters for hand-built tables and for sig == 0.
 * The reference's own kernel, run on a B200 (tests/golden/ref_search_cuckoo_16.npz,
 * dup_* arrays), keeps the LOWEST matching lane's store; that is the rule here
 * and in the CUDA path. */
static inline void search_one(const bucket_t *t, const orc_geom_t *g,
		uint32_t sig, uint32_t hash, uint32_t *o)
{
	const bucket_t *b = &t[orc_bucket1(g, hash)];
	for (int l = SLOTS - 1; l >= 0; l--)
		if (b->sig[l] == sig) o[0] = b->loc[l];
	b = &t[orc_bucket2(g, hash, sig)];        /* always probed, :61-63 is commented out */
	for (int l = SLOTS - 1; l >= 0; l--)
		if (b->sig[l] == sig) o[1] = b->loc[l];
}

void orc_search(const void *table, const orc_geom_t *g,
		const orc_sel_t *in, size_t n, uint32_t *out)
{
	const bucket_t *t = (const bucket_t *)table;
	for (size_t i = 0; i < n; i++)
		search_one(t, g, in[i].sig, in[i].hash, &out[2 * i]);
}

/* The same loop with the two buckets of request i + 12 prefetched while request i is compared: what any CPU
 * implementation that cares about throughput does (the table is far bigger than the caches, every probe is a DRAM
 * miss).  Results are orc_search's word for word (tests/test_oracle.py); used by the thread pool of the CPU baseline so
 * that the GPU is not compared with a latency-bound loop. */
void orc_search_pf(const void *table, const orc_geom_t *g,
		const orc_sel_t *in, size_t n, uint32_t *out)
{
	const bucket_t *t = (const bucket_t *)table;
	const size_t ahead = 12;
	for (size_t i = 0; i < n; i++) {
		if (i + ahead < n) {
			__builtin_prefetch(&t[orc_bucket1(g, in[i + ahead].hash)], 0, 0);
			__builtin_prefetch(&t[orc_bucket2(g, in[i + ahead].hash, in[i + ahead].sig)], 0, 0);
		}
		search_one(t, g, in[i].sig, in[i].hash, &out[2 * i]);
	}
}

/* ------------------------------------------------------------------ insert */

/* lowest lane whose signature equals sig: __ffs(ballot)-1 at gpu_hash.cu:120,282,344 */
static inline int lowest_match(const bucket_t *b, uint32_t sig)
{
	for (int l = 0; l < SLOTS; l++)
		if (b->sig[l] == sig) return l;
	return -1;
}

/* gpu_hash.cu:139-147 / :301-309: the empty-lane ballot is split at the major
 * location m = sig & 7; lanes >= m keep their bit position, lanes < m are moved
 * up by 16, then __ffs picks the lowest: first empty slot in m, m+1, .., 7, 0, .., m-1. */
static inline int first_empty_from(const bucket_t *b, int m)
{
	for (int k = 0; k < SLOTS; k++) {
		int l = (m + k) & (SLOTS - 1);
		if (b->sig[l] == 0) return l;
	}
	return -1;
}

/* gpu_hash.cu:256-430 */
static void insert_cuckoo_one(bucket_t *t, const orc_geom_t *g,
		uint32_t sig0, uint32_t hash, uint32_t loc0, orc_stats_t *st)
{
	if (sig0 == 0 && loc0 == 0) { st->skipped++; return; }          /* :259-262 */

	uint32_t sig = sig0, loc = loc0;
	const int m = (int)(sig0 & (SLOTS - 1));                          /* ml_mask, :301 (never recomputed) */
	bucket_t *b = &t[orc_bucket1(g, hash)];
	int l;

	if ((l = lowest_match(b, sig)) >= 0) {                            /* :277-287 */
		b->loc[l] = loc; st->updated++; st->chain_hist[0]++; return;
	}
	if ((l = first_empty_from(b, m)) >= 0) {                          /* :303-327 */
		b->sig[l] = sig; b->loc[l] = loc; st->placed_b1++; st->chain_hist[0]++; return;
	}

	st->to_b2++;
	int c = 0;                                                        /* cuckoo_num, :331 */
	for (;;) {                                                        /* cuckoo_evict:, :333 */
		/* `hash` is the request's hash even when (sig,loc) is a victim: :334-335, :403-404 */
		b = &t[orc_bucket2(g, hash, sig)];
		if ((l = lowest_match(b, sig)) >= 0) {                        /* :339-349 */
			b->loc[l] = loc; st->updated++; break;
		}
		if ((l = first_empty_from(b, m)) >= 0) {                      /* :351-356, :371-395 */
			b->sig[l] = sig; b->loc[l] = loc; st->placed_b2++; break;
		}
		l = (int)(sig0 & (SLOTS - 1));                                /* elem->sig, :360 */
		if (c < g->max_cuckoo) {                                      /* :361-365, :397-405 */
			uint32_t vs = b->sig[l], vl = b->loc[l];
			b->sig[l] = sig; b->loc[l] = loc;
			c++; st->displaced++;
			sig = vs; loc = vl;
			continue;
		}
		b->sig[l] = sig; b->loc[l] = loc;                             /* :414-422 */
		st->dropped++;
		break;
	}
	st->chain_hist[c < 7 ? c : 7]++;
}

/* gpu_hash.cu:97-226 */
static void insert_2choice_one(bucket_t *t, const orc_geom_t *g,
		uint32_t sig, uint32_t hash, uint32_t loc, orc_stats_t *st)
{
	if (sig == 0 && loc == 0) { st->skipped++; return; }             /* :101-104 */

	const int m = (int)(sig & (SLOTS - 1));                           /* :139 */
	bucket_t *b = &t[orc_bucket1(g, hash)];
	int l;

	if ((l = lowest_match(b, sig)) >= 0) { b->loc[l] = loc; st->updated++; return; }     /* :115-125 */
	if ((l = first_empty_from(b, m)) >= 0) {                                              /* :141-165 */
		b->sig[l] = sig; b->loc[l] = loc; st->placed_b1++; return;
	}
	st->to_b2++;
	b = &t[orc_bucket2(g, hash, sig)];                                                    /* :171-173 */
	if ((l = lowest_match(b, sig)) >= 0) { b->loc[l] = loc; st->updated++; return; }     /* :176-186 */
	if ((l = first_empty_from(b, m)) >= 0) {                                              /* :188-196, :211-220 */
		b->sig[l] = sig; b->loc[l] = loc; st->placed_b2++; return;
	}
	/* :197-209: both buckets full -> the signature replaces slot sig&7 of bucket 2 and the
	 * loop is left through `break` without the loc store: the slot keeps the old location. */
	b->sig[m] = sig;
	st->overwritten++;
}

void orc_insert(void *table, const orc_geom_t *g, const orc_iel_t *in, size_t n, orc_stats_t *st)
{
	orc_stats_t local;
	if (!st) { memset(&local, 0, sizeof local); st = &local; }
	bucket_t *t = (bucket_t *)table;
	if (g->algo == ORC_2CHOICE)                                       /* gpu_hash.cu:537-553 */
		for (size_t i = 0; i < n; i++) insert_2choice_one(t, g, in[i].sig, in[i].hash, in[i].loc, st);
	else
		for (size_t i = 0; i < n; i++) insert_cuckoo_one(t, g, in[i].sig, in[i].hash, in[i].loc, st);
}

/* gpu_hash.cu:82-83, 236-237: CUDA block k consumes blk_input[k][0..blk_elem_num[k]) */
void orc_insert_blocks(void *table, const orc_geom_t *g,
		const orc_iel_t *const *blk, const int *blk_n, int num_blks, orc_stats_t *st)
{
	for (int k = 0; k < num_blks; k++)
		if (blk_n[k] > 0) orc_insert(table, g, blk[k], (size_t)blk_n[k], st);
}

/* ------------------------------------------------------------------ delete */

/* gpu_hash.cu:454-477: every lane with sig AND loc equal zeroes its signature (loc stays);
 * bucket 2 is only visited when no lane of bucket 1 matched (:465-468). */
uint64_t orc_delete(void *table, const orc_geom_t *g, const orc_iel_t *in, size_t n)
{
	bucket_t *t = (bucket_t *)table;
	uint64_t zeroed = 0;
	for (size_t i = 0; i < n; i++) {
		bucket_t *b = &t[orc_bucket1(g, in[i].hash)];
		int hit = 0;
		for (int l = 0; l < SLOTS; l++)
			if (b->sig[l] == in[i].sig && b->loc[l] == in[i].loc) { b->sig[l] = 0; hit++; }
		if (hit) { zeroed += (uint64_t)hit; continue; }
		b = &t[orc_bucket2(g, in[i].hash, in[i].sig)];
		for (int l = 0; l < SLOTS; l++)
			if (b->sig[l] == in[i].sig && b->loc[l] == in[i].loc) { b->sig[l] = 0; zeroed++; }
	}
	return zeroed;
}

/* -------------------------------------------------------------- inspection */

static inline uint64_t mix64(uint64_t z)
{
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

uint64_t orc_table_occupied(const void *table, const orc_geom_t *g)
{
	const bucket_t *t = (const bucket_t *)table;
	uint64_t n = 0, nb = (uint64_t)g->hash_mask + 1;
	for (uint64_t b = 0; b < nb; b++)
		for (int l = 0; l < SLOTS; l++) n += t[b].sig[l] != 0;
	return n;
}

void orc_table_digest(const void *table, const orc_geom_t *g, int per_bucket, uint64_t digest[2])
{
	const bucket_t *t = (const bucket_t *)table;
	uint64_t sum = 0, x = 0, nb = (uint64_t)g->hash_mask + 1;
	for (uint64_t b = 0; b < nb; b++)
		for (int l = 0; l < SLOTS; l++) {
			if (t[b].sig[l] == 0) continue;
			uint64_t h = mix64(((uint64_t)t[b].sig[l] << 32) | t[b].loc[l]);
			if (per_bucket) h = mix64(h ^ (b * 0x9E3779B97F4A7C15ULL));
			sum += h; x ^= h;
		}
	digest[0] = sum; digest[1] = x;
}

/* -------------------------------------------------------------- key stream */

uint64_t orc_splitmix64(uint64_t *state)
{
	uint64_t z = (*state += 0x9E3779B97F4A7C15ULL);
	return mix64(z);
}

/* (sig, hash) from an 8-byte key as the receiver does: src/mega_recv.c:350,361-362 */
void orc_keys_fill(uint64_t seed, uint64_t first_index, size_t n, orc_iel_t *iel, orc_sel_t *sel)
{
	uint64_t state = seed + first_index * 0x9E3779B97F4A7C15ULL;
	for (size_t i = 0; i < n; i++) {
		uint64_t key = orc_splitmix64(&state);
		uint32_t sig = (uint32_t)key, hash = (uint32_t)(key >> 32);
		if (sig == 0) sig = 1;                      /* 0 is the empty marker; cf. src/mega_common.c:58-59 */
		if (iel) { iel[i].sig = sig; iel[i].hash = hash; iel[i].loc = (uint32_t)(first_index + i + 1); }
		if (sel) { sel[i].sig = sig; sel[i].hash = hash; }
	}
}

/* ------------------------------------------------- the reference's zipf key ranks */

/* src/zipf.h:44-71: a^b for 0 <= b: the integer part of b by repeated squaring, the fractional part by scaling the
 * exponent field of the double (high word minus the bias constant 1072632447, low word cleared). */
static double zipf_pow(double a, double b)
{
	int whole = (int)b;
	uint64_t bits; memcpy(&bits, &a, sizeof bits);
	int32_t hi = (int32_t)(bits >> 32);
	hi = (int32_t)((b - (double)whole) * (double)(hi - 1072632447) + 1072632447.);
	bits = (uint64_t)(uint32_t)hi << 32;                      /* :56 low word = 0 */
	double frac; memcpy(&frac, &bits, sizeof frac);
	double r = 1.;
	for (; whole; whole >>= 1, a *= a)                        /* :61-68 */
		if (whole & 1) r *= a;
	return r * frac;
}

/* src/zipf.h:117-126: 48-bit linear congruential step (the drand48 constants), scaled by 1 / (2^48 - 1) */
static double zipf_rand(uint64_t *x)
{
	*x = (*x * 0x5deece66dULL + 0xbULL) & ((1ULL << 48) - 1);
	return (double)*x / (double)((1ULL << 48) - 1);
}

/* src/zipf.h:73-115 (init) and :137-152 (the lazily computed constants): zetan = sum_{i=1..n} 1 / i^theta summed in
 * ascending order with the approximate pow; theta == 0: uniform.  (theta == -1 "sequential" and theta >= 40 "always 0"
 * of the reference are not workloads of this path.) */
void orc_zipf_init(orc_zipf_t *z, uint64_t n, double theta, uint64_t rand_seed)
{
	memset(z, 0, sizeof *z);
	z->n = n; z->theta = theta; z->rand_state = rand_seed; z->dbl_n = (double)n;
	if (theta > 0. && theta < 1.) {
		z->alpha = 1. / (1. - theta);
		z->thres = 1. + zipf_pow(0.5, theta);
		double sum = 0.;
		for (uint64_t i = 0; i < n; i++) sum += 1. / zipf_pow((double)i + 1., theta);          /* :103-115 */
		z->zetan = sum;
		double zeta2 = 0.;
		for (uint64_t i = 0; i < 2; i++) zeta2 += 1. / zipf_pow((double)i + 1., theta);
		z->eta = (1. - zipf_pow(2. / (double)n, 1. - theta)) / (1. - zeta2 / sum);               /* :145-146 */
	}
}

/* a state whose zetan is already known (2^29 terms take seconds; the bench computes it once) */
void orc_zipf_init_zetan(orc_zipf_t *z, uint64_t n, double theta, uint64_t rand_seed, double zetan)
{
	memset(z, 0, sizeof *z);
	z->n = n; z->theta = theta; z->rand_state = rand_seed; z->dbl_n = (double)n;
	z->alpha = 1. / (1. - theta);
	z->thres = 1. + zipf_pow(0.5, theta);
	z->zetan = zetan;
	double zeta2 = 0.;
	for (uint64_t i = 0; i < 2; i++) zeta2 += 1. / zipf_pow((double)i + 1., theta);
	z->eta = (1. - zipf_pow(2. / (double)n, 1. - theta)) / (1. - zeta2 / zetan);
}

/* src/zipf.h:161-182: Gray et al.'s inversion; ranks 0 and 1 by thresholds, the rest by the power law */
uint64_t orc_zipf_next(orc_zipf_t *z)
{
	double u = zipf_rand(&z->rand_state);
	if (z->theta == 0.) return (uint64_t)(z->dbl_n * u);                                       /* :161-165 */
	double uz = u * z->zetan;
	if (uz < 1.) return 0;
	if (uz < z->thres) return 1;
	return (uint64_t)(z->dbl_n * zipf_pow(z->eta * (u - 1.) + 1., z->alpha));
}

void orc_zipf_fill(orc_zipf_t *z, size_t n, uint64_t *ranks)
{
	for (size_t i = 0; i < n; i++) ranks[i] = orc_zipf_next(z);
}

/* ------------------------------------------------------- threaded baseline */

double orc_now_sec(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
	const void *ctable; void *table; const orc_geom_t *g;
	const orc_sel_t *sel; const orc_iel_t *iel; uint32_t *out;
	size_t lo, hi; int part, parts;
} job_t;

static void *search_worker(void *p)
{
	job_t *j = (job_t *)p;
	orc_search(j->ctable, j->g, j->sel + j->lo, j->hi - j->lo, j->out + 2 * j->lo);
	return NULL;
}

void orc_search_mt(const void *table, const orc_geom_t *g,
		const orc_sel_t *in, size_t n, uint32_t *out, int threads)
{
	if (threads < 1) threads = 1;
	pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
	job_t *jobs = (job_t *)calloc((size_t)threads, sizeof(job_t));
	for (int k = 0; k < threads; k++) {
		jobs[k].ctable = table; jobs[k].g = g; jobs[k].sel = in; jobs[k].out = out;
		jobs[k].lo = n * (size_t)k / (size_t)threads;
		jobs[k].hi = n * (size_t)(k + 1) / (size_t)threads;
		pthread_create(&th[k], NULL, search_worker, &jobs[k]);
	}
	for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
	free(jobs); free(th);
}

/* A thread owns the requests whose bucket index has top-3-bit partition id
 * p with p % parts == part, and walks the whole batch in order, so requests of
 * one partition keep their batch order: the result equals orc_insert(). */
static void *insert_worker(void *p)
{
	job_t *j = (job_t *)p;
	const orc_geom_t *g = j->g;
	uint32_t nb = g->hash_mask + 1;
	uint32_t shift = 0;
	while ((nb >> shift) > 8) shift++;              /* bucket >> shift = partition id 0..7 */
	orc_stats_t st; memset(&st, 0, sizeof st);
	for (size_t i = 0; i < j->hi; i++) {
		uint32_t part = (orc_bucket1(g, j->iel[i].hash) >> shift) % (uint32_t)j->parts;
		if ((int)part == j->part) orc_insert(j->table, g, &j->iel[i], 1, &st);
	}
	return NULL;
}

void orc_insert_mt(void *table, const orc_geom_t *g, const orc_iel_t *in, size_t n, int threads)
{
	if (threads < 1) threads = 1;
	if (threads > 8) threads = 8;                   /* IBLOCK_P = 3 closed partitions */
	if ((g->hash_mask + 1) < 8) threads = 1;
	while (8 % threads) threads--;                   /* 1, 2, 4 or 8 */
	pthread_t th[8]; job_t jobs[8];
	memset(jobs, 0, sizeof jobs);
	for (int k = 0; k < threads; k++) {
		jobs[k].table = table; jobs[k].g = g; jobs[k].iel = in; jobs[k].hi = n;
		jobs[k].part = k; jobs[k].parts = threads;
		pthread_create(&th[k], NULL, insert_worker, &jobs[k]);
	}
	for (int k = 0; k < threads; k++) pthread_join(th[k], NULL);
}

/* ---------------------------------------------- persistent pool (CPU baseline) */

/* The CPU arm of bench.py: one scheduler cycle of `batches` worker batches on the host's cores, with threads that live
 * across cycles (one per core, handed work through a barrier) instead of a pthread_create per batch.
 *   phase 1  every thread zeroes its slice of `out` (the caller's memset, mega_scheduler.c:406) and searches its slice
 *            of ALL batches' requests (gpu_hash.cu:28-75) -- the table is read-only in this phase
 *   phase 2  inserts (gpu_hash.cu:231-433 / :77-229): the 8 bucket ranges with equal top IBLOCK_P bits are closed under
 *            the alternate-bucket function (gpu_hash.h:67-69), so range r is owned by thread r % min(threads, 8), which
 *            walks all batches in order -- the table equals orc_insert() over the concatenated batches
 * Per batch that is the reference's in-stream order search -> insert; batches are unordered against each other in the
 * reference (one stream per worker, mega_scheduler.c:276-280), and "all searches, then all inserts" is one such order. */
struct orc_pool_s {
	int threads;
	pthread_t *th;
	pthread_barrier_t start, done;
	int stop;
	/* the job */
	int kind;                              /* 0 cycle, 1 insert only, 2 preload from the key stream */
	void *table; const orc_geom_t *g;
	const orc_sel_t *sel; uint32_t *out; size_t n_search_total;
	const orc_iel_t *iel; size_t n_insert_total;
	uint64_t seed, first, count;
};

typedef struct { struct orc_pool_s *p; int id; } pool_arg_t;

static void pool_insert_range(struct orc_pool_s *p, int id)
{
	const orc_geom_t *g = p->g;
	const int owners = p->threads < 8 ? p->threads : 8;
	if (id >= owners) return;
	uint32_t nb = g->hash_mask + 1, shift = 0;
	if (nb < 8) { if (id == 0) orc_insert(p->table, g, p->iel, p->n_insert_total, NULL); return; }
	while ((nb >> shift) > 8) shift++;
	if (p->kind == 2) {                    /* generate the keys here: every owner walks the stream and keeps its ranges */
		uint64_t state = p->seed + p->first * 0x9E3779B97F4A7C15ULL;
		for (uint64_t i = 0; i < p->count; i++) {
			uint64_t key = orc_splitmix64(&state);
			orc_iel_t e; e.sig = (uint32_t)key; e.hash = (uint32_t)(key >> 32); e.loc = (uint32_t)(p->first + i + 1);
			if (e.sig == 0) e.sig = 1;
			if ((int)((orc_bucket1(g, e.hash) >> shift) % (uint32_t)owners) == id) orc_insert(p->table, g, &e, 1, NULL);
		}
		return;
	}
	for (size_t i = 0; i < p->n_insert_total; i++)
		if ((int)((orc_bucket1(g, p->iel[i].hash) >> shift) % (uint32_t)owners) == id) orc_insert(p->table, g, &p->iel[i], 1, NULL);
}

static void *pool_worker(void *a_)
{
	pool_arg_t *a = (pool_arg_t *)a_;
	struct orc_pool_s *p = a->p;
	const int id = a->id;
	free(a);
	for (;;) {
		pthread_barrier_wait(&p->start);
		if (p->stop) break;
		if (p->kind == 0) {
			size_t lo = p->n_search_total * (size_t)id / (size_t)p->threads, hi = p->n_search_total * (size_t)(id + 1) / (size_t)p->threads;
			memset(p->out + 2 * lo, 0, (hi - lo) * 2 * sizeof(uint32_t));
			orc_search_pf(p->table, p->g, p->sel + lo, hi - lo, p->out + 2 * lo);
			pthread_barrier_wait(&p->done);          /* all searches of the cycle before any insert */
		}
		pool_insert_range(p, id);
		pthread_barrier_wait(&p->done);
	}
	return NULL;
}

orc_pool_t *orc_pool_create(int threads)
{
	if (threads < 1) threads = 1;
	struct orc_pool_s *p = (struct orc_pool_s *)calloc(1, sizeof *p);
	p->threads = threads;
	p->th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
	pthread_barrier_init(&p->start, NULL, (unsigned)threads + 1);
	pthread_barrier_init(&p->done, NULL, (unsigned)threads + 1);
	for (int k = 0; k < threads; k++) {
		pool_arg_t *a = (pool_arg_t *)malloc(sizeof *a); a->p = p; a->id = k;
		pthread_create(&p->th[k], NULL, pool_worker, a);
	}
	return p;
}

int orc_pool_threads(const orc_pool_t *p) { return p->threads; }

void orc_pool_destroy(orc_pool_t *p)
{
	p->stop = 1;
	pthread_barrier_wait(&p->start);
	for (int k = 0; k < p->threads; k++) pthread_join(p->th[k], NULL);
	pthread_barrier_destroy(&p->start); pthread_barrier_destroy(&p->done);
	free(p->th); free(p);
}

static void pool_run(orc_pool_t *p)
{
	pthread_barrier_wait(&p->start);
	if (p->kind == 0) pthread_barrier_wait(&p->done);
	pthread_barrier_wait(&p->done);
}

/* sel / out / iel hold the `batches` batches of the cycle back to back (n_search, 2 n_search, n_insert entries each) */
void orc_pool_cycle(orc_pool_t *p, void *table, const orc_geom_t *g, const orc_sel_t *sel, size_t n_search, uint32_t *out,
		const orc_iel_t *iel, size_t n_insert, int batches)
{
	p->kind = 0; p->table = table; p->g = g;
	p->sel = sel; p->out = out; p->n_search_total = n_search * (size_t)batches;
	p->iel = iel; p->n_insert_total = n_insert * (size_t)batches;
	pool_run(p);
}

void orc_pool_insert(orc_pool_t *p, void *table, const orc_geom_t *g, const orc_iel_t *iel, size_t n)
{
	p->kind = 1; p->table = table; p->g = g; p->iel = iel; p->n_insert_total = n;
	pool_run(p);
}

/* keys first .. first+count-1 of the SURVEY 8(d) stream inserted without materialising them */
void orc_pool_preload(orc_pool_t *p, void *table, const orc_geom_t *g, uint64_t seed, uint64_t first, uint64_t count)
{
	p->kind = 2; p->table = table; p->g = g; p->seed = seed; p->first = first; p->count = count;
	pool_run(p);
}

/* n searches for keys drawn uniformly from the first `population` keys of the stream (every one hits once they are in) */
void orc_gen_queries(uint64_t seed, uint64_t population, size_t n, uint64_t rng_seed, orc_sel_t *out)
{
	uint64_t r = rng_seed * 0xD1342543DE82EF95ULL + 0x2545F4914F6CDD1DULL;
	for (size_t i = 0; i < n; i++) {
		uint64_t idx = (uint64_t)(((__uint128_t)orc_splitmix64(&r) * population) >> 64);
		uint64_t state = seed + idx * 0x9E3779B97F4A7C15ULL;
		uint64_t key = orc_splitmix64(&state);
		out[i].sig = (uint32_t)key; out[i].hash = (uint32_t)(key >> 32);
		if (out[i].sig == 0) out[i].sig = 1;
	}
}
