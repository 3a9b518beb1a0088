/*
 * gpu_hash.h -- table geometry and request types of the Mega-KV GPU hash index.
 *
 * Drop-in replacement for the reference header of the same name
 * (reference: libgpuhash/gpu_hash.h:38-104).  Every macro and type name that
 * a caller of libgpuhash consumes (mega.c:132-133,290; mega_scheduler.c:273,482;
 * mega_recv.c:571,659; libgpuhash/test/ *.c) is kept with the same value and the
 * same memory layout; what differs:
 *
 *   - MEM_P and the placement policy can be set from the compiler command line
 *     (-DMEM_P=34 -DHASH_2CHOICE); without flags the reference defaults hold
 *     (MEM_P 30, HASH_CUCKOO).
 *   - the size macros are computed in 64-bit arithmetic, so MEM_P >= 31 does
 *     not overflow `int` (reference gpu_hash.h:64 does).
 *   - the unused selem_shared_t / UNIT_THREAD_NUM leftovers are still declared
 *     so that old code keeps compiling.
 *
 *   - INSIDE A BUCKET the library is free to order the 16 words differently
 *     from bucket_t: the reference's callers only allocate and zero-fill the
 *     table (mega_scheduler.c:273-274) and never read or write its bytes, and
 *     all-zero means empty in every layout.  By default the library keeps slot
 *     l as the pair {sig, loc} at words 2l, 2l+1 (one 64-bit CAS per commit)
 *     instead of bucket_t's sig[8] then loc[8].  A host that DOES build or read
 *     table images through bucket_t (libgpuhash/test/back/py_search_stream.c:
 *     104-121) must select the reference byte layout: build the library with
 *     -DGPUHASH_DEFAULT_LAYOUT_REFERENCE, or call gpuhash_set_default_geom()
 *     with layout = GPUHASH_LAYOUT_REFERENCE, or convert with
 *     gpuhash_table_convert() / gpuhash_index_load() / gpuhash_index_dump().
 *   - the caller's MEM_P must be the library's: the three legacy calls take no
 *     size, so a table smaller than the library's HT_SIZE is written out of
 *     bounds exactly as with the reference (whose MEM_P is compiled in too).
 *
 * The values baked in here are only the *defaults* of the three legacy entry
 * points in libgpuhash.h; the extended API (gpuhash_ex.h) takes the geometry
 * at run time.
 */
#ifndef _GPU_HASH_H_
#define _GPU_HASH_H_

#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#  define MY_ALIGN(n) __align__(n)
#elif defined(__GNUC__)
#  define MY_ALIGN(n) __attribute__((aligned(n)))
#elif defined(_MSC_VER)
#  define MY_ALIGN(n) __declspec(align(n))
#else
#  error "MY_ALIGN: unknown host compiler"
#endif

/* 32-bit key signature, 32-bit item location, 32-bit bucket hash. */
typedef uint32_t sign_t;
typedef uint32_t loc_t;
typedef uint32_t hash_t;

/* ---- bucket geometry (reference gpu_hash.h:46-51) ---- */
#define ELEM_SIG_SIZE      8
#define ELEM_SIZE_P        3                       /* log2(sizeof sig + sizeof loc)  */
#define ELEM_NUM_P         3                       /* log2(slots per bucket)         */
#define ELEM_NUM           (1 << ELEM_NUM_P)       /* 8 slots                         */
#define UNIT_THREAD_NUM_P  1
#define UNIT_THREAD_NUM    (1 << UNIT_THREAD_NUM_P)

/* ---- table size: 2^MEM_P bytes (reference gpu_hash.h:54-64) ---- */
#ifndef MEM_P
#  define MEM_P            (30)
#endif
#define BUC_P              (ELEM_NUM_P + ELEM_SIZE_P)               /* 64 B per bucket */
#define BUC_NUM            (1 << (MEM_P - BUC_P))                   /* int up to MEM_P 36 */
#define HASH_MASK          ((1 << (MEM_P - BUC_P)) - 1)
#define HT_SIZE            ((size_t)1 << (MEM_P))                   /* 64-bit: MEM_P >= 31 is fine */

/* ---- insert partitions: the top IBLOCK_P bits of a bucket index are shared by
 * both candidate buckets of a key (reference gpu_hash.h:66-69) ---- */
#define IBLOCK_P           3
#define INSERT_BLOCK       (1 << IBLOCK_P)
#define BLOCK_HASH_MASK    ((1 << (MEM_P - BUC_P - IBLOCK_P)) - 1)

/* ---- placement policy (reference gpu_hash.h:72-76): exactly one is active ---- */
#if defined(HASH_2CHOICE) && defined(HASH_CUCKOO)
#  error "define only one of HASH_2CHOICE / HASH_CUCKOO"
#endif
#if !defined(HASH_2CHOICE) && !defined(HASH_CUCKOO)
#  define HASH_CUCKOO      1
#endif
#ifdef HASH_CUCKOO
#  define MAX_CUCKOO_NUM   5      /* displacements before a victim is dropped */
#endif

/* One bucket = one 32 B signature row followed by one 32 B location row
 * (reference gpu_hash.h:79-82).  sizeof == 64, array stride 64. */
typedef MY_ALIGN(128) struct bucket_s {
	sign_t sig[ELEM_NUM];
	loc_t  loc[ELEM_NUM];
} bucket_t;

/* search request (reference gpu_hash.h:85-89) */
typedef MY_ALIGN(8) struct selem_s {
	sign_t sig;
	hash_t hash;
} selem_t;

typedef MY_ALIGN(8) union selem_shared_s {
	selem_t   elem_s;
	long long elem_u;
} selem_shared_t;

/* insert / delete request, 12 B packed (reference gpu_hash.h:98-104) */
typedef struct ielem_s {
	sign_t sig;
	hash_t hash;
	loc_t  loc;
} ielem_t;

typedef struct ielem_s delem_t;

#endif /* _GPU_HASH_H_ */
