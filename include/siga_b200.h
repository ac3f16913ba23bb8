/*
 * siga_b200 -- C ABI of the B200-native SIGA hot path (`siga index` + `siga overlap`).
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, no C++/torch types.  Every entry point
 * names the reference interface it replaces (paths relative to the chungongyu/siga tree).  The reference has no
 * FFI of its own; the seams are SuffixArrayBuilder::build (+ BWT ctor), FMIndex::load/getOcc
 * and the body of parallel::foreach in OverlapBuilder::build.  INTEGRATION.md shows the C++ stubs a maintainer adds.
 *
 * Conventions
 *   - every function returns 0 on success, a negative SGX_E* code on failure; sgx_last_error() gives the text.
 *   - there is NO CPU fallback: if no CUDA device / kernel image is usable, sgx_init fails with SGX_ENODEV.
 *   - one sgx_ctx per GPU (one process or thread per GPU); calls on one ctx are not re-entrant.
 *   - inputs are borrowed for the duration of the call; outputs are owned by the library (pinned host memory)
 *     and released with the matching *_free.
 *   - symbol ranks follow src/alphabet.h:13-39: '$'=0 'A'=1 'C'=2 'G'=3 'T'=4.
 */
#ifndef SIGA_B200_H_
#define SIGA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGX_ABI_VERSION 2

enum {
  SGX_OK = 0,
  SGX_EINVAL = -1,   /* bad argument (null pointer, empty read, non-ACGT base, ...)        */
  SGX_ENODEV = -2,   /* no usable CUDA device / sm_100a kernel image -- never falls back    */
  SGX_ECUDA = -3,    /* a CUDA runtime call or kernel failed                                */
  SGX_ENOMEM = -4,   /* device or pinned-host allocation failed                             */
  SGX_EOVERFLOW = -5 /* an on-device work area overflowed even after the regrow-and-retry   */
};

/* Suffix ordering models of the two reference builders. */
enum {
  SGX_ORDER_SAIS = 0, /* `-a sais`  (SAISBuilder, src/suffix_array_builder.cpp:31-282): ties by read index   */
  SGX_ORDER_SAIS2 = 1 /* `-a sais2` (NewSAISBuilder, src/suffix_array_builder.cpp:474-673, the default):
                         plain suffix array of R0$R1$...$, all '$' equal, end of text smallest                */
};

/* AlignFlags bits, src/overlap_builder.cpp:42-44. */
enum { SGX_AF_QUERYREV = 1, SGX_AF_TARGETREV = 2, SGX_AF_QUERYCOMP = 4 };

typedef struct sgx_ctx sgx_ctx;
typedef struct sgx_reads sgx_reads;
typedef struct sgx_fm sgx_fm;

/* ---- context ------------------------------------------------------------------------------------------ */

int sgx_abi_version(void);
/* Create a context on CUDA device `device`.  Replaces nothing in the reference (it has no device). */
int sgx_init(int device, sgx_ctx** out);
void sgx_shutdown(sgx_ctx* ctx);
/* Text of the last error on this context (or of the last failed sgx_init when ctx == NULL). */
const char* sgx_last_error(const sgx_ctx* ctx);
/* Streaming-multiprocessor count of the context's device. */
int sgx_sm_count(const sgx_ctx* ctx);

/* ---- reads -------------------------------------------------------------------------------------------- */

/*
 * Pack `n` reads into the device read store (2 bits/base + lengths).  `ascii` holds the bases of all reads
 * back to back, read i occupying [off[i], off[i+1]).  Replaces the in-memory DNASeqList produced by
 * ReadDNASequences (src/kseq.cpp:230-257) as the input of both hot paths.
 * Deviation (documented): bytes other than A/C/G/T and empty reads are rejected with
 * SGX_EINVAL instead of being mapped to rank 0.
 */
int sgx_reads_pack(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n, sgx_reads** out);
/*
 * sgx_reads_pack without the wait: returns once the read lengths are known (one small upload); the bases cross PCIe in
 * chunks of ~32 MB on a second stream of the context and are packed as they land, while the caller goes on (host ->
 * device streaming).  sgx_overlap_edges starts each of its search batches as soon as the
 * chunks holding that batch's query reads have arrived; every other consumer waits for the whole store first.
 * `ascii` must stay valid, and a byte other than A/C/G/T is reported (SGX_EINVAL), by the first call that consumes the
 * store or by sgx_reads_wait.
 */
int sgx_reads_pack_async(sgx_ctx* ctx, const char* ascii, const uint64_t* off, uint64_t n, sgx_reads** out);
int sgx_reads_wait(sgx_ctx* ctx, const sgx_reads* reads);
void sgx_reads_free(sgx_reads* reads);
uint64_t sgx_reads_count(const sgx_reads* reads);
uint64_t sgx_reads_bases(const sgx_reads* reads);

/* ---- index construction (`siga index`) ------------------------------------------------------------------ */

typedef struct sgx_index_out {
  uint64_t n_strings;  /* reads                                   (.bwt header field 2, src/bwt.cpp:139-150)   */
  uint64_t n_suffixes; /* bases + reads = BWT length              (.bwt header field 3)                        */
  uint64_t n_runs;     /* run bytes                               (.bwt header field 4)                        */
  uint8_t* rle;        /* n_runs bytes, (rank<<5)|count, greedy-31 (src/rlstring.h:10-13, src/bwt.cpp:16-27)   */
  uint32_t* sai;       /* n_strings read ids in the order of the read-start suffixes = the `<i> 0` lines of the
                          .sai file (src/suffix_array.cpp:16-43)                                               */
  double ms_total;     /* device time of the whole build (CUDA events on the ctx stream)                       */
  double ms_sort;      /* read ordering (sentinel order / .sai)                                                */
  double ms_bcr;       /* column-insertion passes                                                              */
  double ms_emit;      /* run-length encoding + copies out                                                     */
  uint64_t launches;   /* kernels launched by this call                                                        */
  uint64_t merge_launches; /* launches of the dominant (merge) kernel                                          */
  double ms_merge;     /* device time spent in the merge kernel                                                */
  uint64_t merge_bytes;/* bytes it read + wrote (algorithmic traffic of the merge passes)                      */
  uint64_t seg_offset; /* sgx_index_build_sharded_segments: `rle` holds seg_bytes run bytes that belong at byte   */
  uint64_t seg_bytes;  /* seg_offset of the n_runs run bytes of the whole index; elsewhere 0 and n_runs          */
} sgx_index_out;

/*
 * Build the BWT and the lexicographic read index of `reads` (or of the reversed reads when reversed != 0,
 * the reference's .rbwt/.rsai, src/indexer.cpp:60-64).  Replaces SuffixArrayBuilder::create(name)->build()
 * (src/suffix_array_builder.h:10-16, .cpp:684-693) + BWT::BWT(sa, reads) (src/bwt.cpp:7-32): the caller writes
 * the 30-byte header + rle as .bwt and the sai list as .sai and obtains files byte-identical to the reference.
 */
int sgx_index_build(sgx_ctx* ctx, const sgx_reads* reads, int order, int reversed, sgx_index_out* out);
void sgx_index_out_free(sgx_index_out* out);

/*
 * Both indices of a read set in one call -- what one `siga index` run produces (Indexer::run builds the index of the reads
 * and then of the reversed reads, src/indexer.cpp:25-77) -- built concurrently on two streams of the device.  The results
 * are those of sgx_index_build(.., reversed = 0, ..) and sgx_index_build(.., reversed = 1, ..); release both with
 * sgx_index_out_free.  The time fields of each result cover that build alone; the two overlap in time.
 */
int sgx_index_build_pair(sgx_ctx* ctx, const sgx_reads* reads, int order, sgx_index_out* out_fwd, sgx_index_out* out_rev);

/* ---- multi-GPU index construction ------------------------------------------------------------------------ */

/*
 * Shard communicator.  The reference has no counterpart (it is a single-node OpenMP program); this is the
 * plumbing of BASELINE.json's multi-GPU index build.  Two transports:
 *   sgx_comm_init        one process per GPU, NCCL underneath (libnccl.so.2 is resolved at run time).  Rank 0 calls
 *                        sgx_comm_unique_id, the caller's launcher (torch.distributed, MPI, a file) carries the 128
 *                        bytes to every rank, every rank calls sgx_comm_init with the ctx of its own GPU.
 *   sgx_comm_init_local  `shards` virtual shards inside one context: same algorithm, device copies instead of NCCL
 *                        (used to test the sharded build on one GPU).
 */
typedef struct sgx_comm sgx_comm;
int sgx_comm_unique_id(uint8_t id[128]);
int sgx_comm_init(sgx_ctx* ctx, const uint8_t id[128], int rank, int world, sgx_comm** out);
int sgx_comm_init_local(sgx_ctx* ctx, int shards, sgx_comm** out);
void sgx_comm_free(sgx_comm* comm);
int sgx_comm_world(const sgx_comm* comm);
int sgx_comm_rank(const sgx_comm* comm);

/*
 * sgx_index_build with the rows of the BWT sharded over the communicator by leading k-mer bucket: every pass the
 * shards exchange their inserts (all-to-all), at the end the run bytes are concatenated in bucket order.  Collective:
 * every rank calls it with the same (replicated) read set.  The result is byte-identical to sgx_index_build;
 * `out->rle` is complete on rank 0 only (other ranks get n_runs and an empty buffer), `out->sai` on every rank.
 */
int sgx_index_build_sharded(sgx_ctx* ctx, sgx_comm* comm, const sgx_reads* reads, int order, int reversed,
                            sgx_index_out* out);
/*
 * Same build without the final concatenation on rank 0: every rank gets the run bytes of its own shard
 * (`out->rle`, `out->seg_bytes` bytes) and the place they take in the index (`out->seg_offset`; `out->n_runs` is the
 * length of the whole).  The segments of ranks 0, 1, ... laid end to end are the run bytes sgx_index_build returns, so
 * the ranks can write one .bwt file together (header by rank 0, every segment with pwrite at 30 + seg_offset) and
 * nothing is gathered.  `out->sai` on every rank.  With a local communicator the one process holds all shards and the
 * result is that of sgx_index_build_sharded.
 */
int sgx_index_build_sharded_segments(sgx_ctx* ctx, sgx_comm* comm, const sgx_reads* reads, int order, int reversed,
                                     sgx_index_out* out);

/* ---- FM-index (`FMIndex::load`, rank queries) ----------------------------------------------------------- */

/*
 * Build the device rank structure from the run bytes of a .bwt file.  Replaces FMIndex::load / initialize
 * (src/fmindex.cpp:124-161, 347-366): the reference's large/small markers are never stored on disk, so the
 * in-memory layout is free; ours is one 128-byte block per 256 symbols.
 */
int sgx_fm_create(sgx_ctx* ctx, const uint8_t* rle, uint64_t n_runs, uint64_t n_suffixes, sgx_fm** out);
void sgx_fm_free(sgx_fm* fm);
/*
 * Attach the read-start order of the index (the ids of its .sai / .rsai file, SuffixArray::load,
 * src/suffix_array.cpp:103-118) and the indexed read set.  The reference loads both for the hits -> ED step
 * (src/overlap_builder.cpp:464-480); with them on the device the irreducible extraction reads the extension
 * symbols of a candidate from the target read itself instead of walking the BWT one rank query per base
 * (identical blocks, DESIGN.md "post stage").  Without it sgx_overlap_batch uses the BWT-walking path.
 * `indexed_reads` is borrowed and must outlive the fm object.
 */
int sgx_fm_attach_reads(sgx_ctx* ctx, sgx_fm* fm, const uint32_t* sai, uint64_t n_strings, const sgx_reads* indexed_reads);
uint64_t sgx_fm_length(const sgx_fm* fm);
/* C[] of FMIndex::getPC (src/fmindex.h:129-131): pred[c] = number of symbols with rank < c. */
int sgx_fm_pred(const sgx_fm* fm, uint64_t pred[5]);
/*
 * Batched FMIndex::getOcc(i) (src/fmindex.cpp:320-323): counts[5*k..5*k+4] = occurrences of $,A,C,G,T in
 * BWT[0..pos[k]] inclusive; pos == UINT64_MAX gives zeros exactly as the reference's wrap-around does.
 */
int sgx_fm_occ(sgx_ctx* ctx, const sgx_fm* fm, const uint64_t* pos, uint64_t n, uint64_t* counts);
/*
 * Batched FMIndex::Interval::occurrences (src/fmindex.h:67-86): number of occurrences of each query string
 * (queries given like sgx_reads_pack's input).  Used by `correct`/`match`/`preqc` style callers.
 */
int sgx_fm_count(sgx_ctx* ctx, const sgx_fm* fm, const char* ascii, const uint64_t* off, uint64_t n, uint64_t* counts);

/* ---- overlap search (`siga overlap`) -------------------------------------------------------------------- */

/* One OverlapBlock (src/overlap_builder.cpp:151-196) exactly as the reference serialises it into .hits. */
typedef struct sgx_block {
  uint64_t capped[4]; /* capped[0].lower, capped[0].upper, capped[1].lower, capped[1].upper */
  uint64_t raw[4];    /* raw[0].lower,    raw[0].upper,    raw[1].lower,    raw[1].upper    */
  uint32_t length;    /* overlap length                                                     */
  uint32_t af;        /* SGX_AF_* bits                                                      */
} sgx_block;

typedef struct sgx_hits {
  uint64_t n_reads;
  uint8_t* substring;   /* n_reads flags: OverlapResult.substring (src/overlap_builder.cpp:211-216)            */
  uint64_t* block_off;  /* n_reads + 1 offsets into blocks                                                     */
  sgx_block* blocks;    /* blocks of read i: [block_off[i], block_off[i+1]) (order within a read is free)      */
  uint64_t n_blocks;
  double ms_total;      /* device time of the call incl. copies (CUDA events)                                  */
  double ms_search;     /* the four-way backward search kernel                                                 */
  double ms_extract;    /* sub-maximal filter + irreducible extraction kernel                                  */
  uint64_t lookups_search;  /* executed 128-byte rank-block fetches in the search kernel                       */
  uint64_t lookups_extract; /* executed rank-block fetches in filter/extraction                                */
  uint64_t launches;    /* kernels launched by this call                                                       */
  uint64_t search_launches;
  uint64_t slow_reads;  /* reads whose post stage took the general BWT-walking path                            */
  uint64_t lines_search; /* 128-byte lines those fetches touched (one when both interval ends share a rank block) */
} sgx_hits;

/*
 * Overlap blocks of reads [first, first+count) of `reads` against the forward and reverse FM-indices.
 * Replaces the body of parallel::foreach in OverlapBuilder::build (src/overlap_builder.cpp:453-457), i.e.
 * OverlapBuilder::overlap (src/overlap_builder.cpp:1113-1182) per read: four OverlapBlockFinder::find passes,
 * SubMaximalBlockFilter, ContainmentBlockRemover, and (irreducible != 0) IrreducibleBlockListExtractor.
 * rc != 0 searches the reverse complement too (the default; `--no-opposite-strand` clears it).
 */
int sgx_overlap_batch(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                      uint64_t count, uint32_t min_overlap, int irreducible, int rc, sgx_hits* out);
void sgx_hits_free(sgx_hits* hits);

/*
 * Same, with the reads given in HOST memory (end-to-end entry: pack + H2D + search + D2H inside the call).
 */
int sgx_overlap_host(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const char* ascii, const uint64_t* off,
                     uint64_t n, uint32_t min_overlap, int irreducible, int rc, sgx_hits* out);

/*
 * One ASQG edge before formatting: Hit2OverlapConverter::convert + OverlapBlock::overlap
 * (src/overlap_builder.cpp:158-175, 345-375) reduced to what the ED line is computed from.  With lq/lt the lengths
 * of query/target:  c1 = [lq-length, lq-1], flipped when af & SGX_AF_QUERYREV;  c2 = [0, length-1], flipped when
 * af & SGX_AF_TARGETREV;  isRC = (af & SGX_AF_QUERYCOMP) != 0;  line = "ED\t<q> <t> c1.s c1.e lq c2.s c2.e lt isRC 0".
 */
typedef struct sgx_edge {
  uint32_t query;   /* index of the query read  (id0) */
  uint32_t target;  /* index of the target read (id1) */
  uint32_t length;  /* overlap length                 */
  uint32_t af;      /* SGX_AF_* bits                  */
} sgx_edge;

typedef struct sgx_edges {
  uint64_t n_reads;
  uint8_t* substring;  /* n_reads flags (VT record SS:i tag)                                        */
  uint64_t n_edges;
  sgx_edge* edges;     /* grouped by query read, order within a read free                           */
  uint64_t n_blocks;   /* overlap blocks the edges were expanded from                               */
  double ms_total, ms_search, ms_extract;
  uint64_t lookups_search, lookups_extract, launches, search_launches, slow_reads;
  uint64_t lines_search; /* distinct 128-byte rank-block lines the search fetched: a backward step whose interval ends
                            share a block counts one line for its two lookups (the random-access roofline numerator) */
} sgx_edges;

/*
 * sgx_overlap_batch followed by the reference's hits -> ED conversion, all on the device: replaces the
 * parallel::foreach body AND the Hits2ASQGConverter pass over the .hits.gz files (src/overlap_builder.cpp:453-480),
 * which the reference runs serially through text files.  Needs sgx_fm_attach_reads on both indices; `reads` must
 * be the indexed read set.  name_rank[i] = rank of read i's name in std::string order (equal names, equal ranks):
 * the reference drops self hits by name and keeps an edge only when id0 is not smaller than id1
 * (src/overlap_builder.cpp:358-367).
 */
int sgx_overlap_edges(sgx_ctx* ctx, const sgx_fm* fwd, const sgx_fm* rev, const sgx_reads* reads, uint64_t first,
                      uint64_t count, uint32_t min_overlap, int irreducible, int rc, const uint32_t* name_rank,
                      sgx_edges* out);
void sgx_edges_free(sgx_edges* edges);

/*
 * Final gather of a read-sharded overlap run ("no collective beyond a final gather of ASQG
 * records"): rank r searched the contiguous read range r of `world` ranges with sgx_overlap_edges; on `root` the
 * result holds the SS flags of all reads in input order (rank order) and every rank's edge records; other ranks get
 * an empty result.  Collective over an NCCL communicator (sgx_comm_init).  The ED order is free in the reference too
 * (thread order, src/overlap_builder.cpp:447-480); VT order = input order is preserved by the rank-ordered flags.
 */
int sgx_edges_gather(sgx_ctx* ctx, sgx_comm* comm, const sgx_edges* local, int root, sgx_edges* out);

/* ---- k-mer error correction (`siga correct`) -------------------------------------------------------------- */

/* Options of KmerCorrector (src/correct_processor.cpp:74-79; defaults src/correct_processor.h:15-20). */
typedef struct sgx_correct_params {
  uint32_t kmer_size;         /* -k, default 31                                                                */
  int32_t kmer_threshold;     /* -x, default 3: a k-mer seen fewer times is a correction candidate             */
  uint32_t kmer_rounds;       /* -i, default 10 (values above 2^20 are treated as 2^20)                        */
  uint64_t kmer_count_offset; /* -O, default 1                                                                 */
} sgx_correct_params;

/* One substituted base of a written read: position = pos_code >> 2, new base = "ACGT"[pos_code & 3]. */
typedef struct sgx_fix {
  uint32_t read;     /* index relative to `first` */
  uint32_t pos_code;
} sgx_fix;

typedef struct sgx_corrected {
  uint64_t n_reads;
  uint8_t* valid;      /* n_reads flags: CorrectResult.validQC -- 1: every base ends covered by a solid k-mer and the
                          read is written (with its fixes); 0: PostCorrector drops it (src/correct_processor.cpp:258-264) */
  uint64_t n_fixes;
  sgx_fix* fixes;      /* substitutions of the written reads, any order                                         */
  double ms_total;     /* device time of the call incl. copies (CUDA events)                                    */
  double ms_kernel;    /* the correction kernel                                                                 */
  uint64_t lookups;    /* executed 128-byte rank lookups (two per backward-search step, as in the overlap search) */
  uint64_t fixed_reads;
  uint64_t launches;
} sgx_corrected;

/*
 * KmerCorrector::process (src/correct_processor.cpp:72-232) for reads [first, first+count) of `reads` against the
 * FORWARD index `fm` (the reference loads PREFIX.bwt only, src/correct.cpp:45-46); replaces the body of
 * parallel::foreach in CorrectProcessor::process (src/correct_processor.cpp:282-287).  qual / qual_off: the FASTQ quality
 * strings of exactly those reads in HOST memory, laid out like sgx_reads_pack's input (qual_off has count + 1
 * entries); both NULL for FASTA input (DNASeq::score then returns 15, src/kseq.h:34-40).
 * SGX_EINVAL: kmer_size 0 (the reference asserts), a read longer than 32768 bases, a quality string whose length
 * differs from its read's.  Reads shorter than kmer_size are dropped (valid = 0), as in the reference.
 */
int sgx_correct_batch(sgx_ctx* ctx, const sgx_fm* fm, const sgx_reads* reads, uint64_t first, uint64_t count,
                      const char* qual, const uint64_t* qual_off, const sgx_correct_params* params, sgx_corrected* out);
void sgx_corrected_free(sgx_corrected* out);

/* ---- measurement helpers --------------------------------------------------------------------------------- */

/*
 * Random 128-byte gather micro-benchmark: `n_loads` independent uniformly random
 * 128-byte-aligned block fetches over a `bytes`-sized buffer; returns fetches per second.
 */
int sgx_bench_gather(sgx_ctx* ctx, uint64_t bytes, uint64_t n_loads, double* fetches_per_s);
/* Write > L2-size bytes so that the next timed kernel starts with a cold L2. */
int sgx_flush_l2(sgx_ctx* ctx);

#ifdef __cplusplus
}
#endif

#endif /* SIGA_B200_H_ */
