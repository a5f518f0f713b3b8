/* bella_xdrop.h -- C-ABI of the "next" row f1: batched gapped X-drop seed-and-extend on one B200.
 *
 * Stands for the alignment step BELLA runs on every nonzero of the overlap matrix right after the SpGEMM:
 *   CPU build:  alignSeqAn(row, col, rowLen, i, j, xDrop, kmerSize)          include/align.hpp:93-139
 *               -> seqan::extendSeed(seed, H, V, EXTEND_BOTH, scoring(1,-1,-1), xDrop, GappedXDrop())
 *                                                                             seqan/seqan/seeds/seeds_extension.h:622-843
 *               followed by PostAlignDecision(...)                           include/overlap.hpp:415-462
 *   GPU build:  alignLogan(target, query, seeds, bpars, results)             include/align.hpp:210-255
 *               -> extendSeedL(...) / extendSeedLGappedXDropOneDirection     loganGPU/functions.cuh:223-689
 *               called from RunPairWiseAlignmentsGPU                         include/overlap.hpp:876-1061
 * Per pair the result is what seqAnResult / loganResult carry -- score, strand, the extended seed (begH, endH, begV, endV;
 * on the reverse strand the H coordinates are on the reverse complement of the row read, as in the reference) -- plus the
 * estimated overlap and the pass/fail of the reference's adaptive threshold.  Bit-exact with the CPU path.
 *
 * Plain pointers and sizes; every function returns 0 or a negative BELLA_XDROP_E* code; bella_xdrop_last_error() explains.
 * Implemented by bella_b200/libbella_xdrop.so (bella_b200/csrc/bella_xdrop.cu + xdrop.cuh).  There is no CPU fallback.
 */
#ifndef BELLA_XDROP_H
#define BELLA_XDROP_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BELLA_XDROP_OK        0
#define BELLA_XDROP_EINVAL   (-1)
#define BELLA_XDROP_ECUDA    (-2)
#define BELLA_XDROP_ESEED    (-3)    /* a seed k-mer does not fit inside its read */

#define BELLA_XDROP_OUT_FIELDS 8     /* score, strand ('n' = 110, 'c' = 99), begH, endH, begV, endV, ov, passed */

typedef struct bella_xdrop bella_xdrop;

bella_xdrop* bella_xdrop_create(int device);
void         bella_xdrop_destroy(bella_xdrop* h);
const char*  bella_xdrop_last_error(const bella_xdrop* h);

/* The read set (readVector_ of the reference: reads[i].seq), concatenated, one byte per base; read i is
 * seqs[seq_off[i] .. seq_off[i+1]).  Copied to the device once and kept there for every later batch (the reference
 * copies both reads of every pair per batch, loganGPU/functions.cuh:470-560). */
int bella_xdrop_set_reads(bella_xdrop* h, const char* seqs, const uint64_t* seq_off, uint32_t n_reads);

/* Scoring is the reference's fixed scheme (match 1, mismatch -1, gap -1: align.hpp:98, :217).  xdrop = BELLApars.xDrop,
 * kmer_len = BELLApars.kmerSize; ratiophi / delta_chernoff / fixed_threshold (-1 = adaptive) are the arguments of
 * PostAlignDecision (overlap.hpp:415-455). */
int bella_xdrop_set_params(bella_xdrop* h, int kmer_len, int xdrop, double ratiophi, double delta_chernoff, int fixed_threshold);

/* Lanes per extension and cells (window slots) per lane: (32,1) (32,2) (32,4) (16,1) (16,2) (16,4) (8,4) (8,8) = the
 * register kernel; (1,32) (1,64) = one thread per extension with the window in shared memory; (2,32) (2,64) = the same
 * with one packed word per cell, (3,W) = packed + longest extension first, (4,W) = both anti-diagonals of a column in one
 * word (4 W bytes per thread; W up to 256 for larger x), (5,W) = that + longest first (2..5 are opt-in: checked under the CPU emulator, not measured yet);
 * (0,0) = everything through the wide path; (-1,-1) = chosen from xdrop (default). */
int bella_xdrop_set_shape(bella_xdrop* h, int lanes, int cells_per_lane);

/* One batch of candidate pairs: rows[p] = H read (row of C), cols[p] = V read (column of C), posH/posV = the seed k-mer the
 * overlap SpGEMM chose (bella_b200_numeric's rowids / posH / posV with the column index expanded).  Host arrays in,
 * out = int32 [n_pairs][8] on the host. */
int bella_xdrop_align(bella_xdrop* h, uint64_t n_pairs, const uint32_t* rows, const uint32_t* cols,
                      const uint16_t* posH, const uint16_t* posV, int32_t* out);

/* Same with device pointers in and out (seeds taken straight from the SpGEMM's device result); asynchronous on the
 * handle's stream. */
int bella_xdrop_align_device(bella_xdrop* h, uint64_t n_pairs, const uint32_t* d_rows, const uint32_t* d_cols,
                             const uint16_t* d_posH, const uint16_t* d_posV, int32_t* d_out);

/* Same, fed with the overlap SpGEMM's result exactly as bella_b200_result_device() hands it out (include/bella_b200.h):
 * C in CSC form over all n_cols = n_reads columns, pair p = nonzero p, its V read = the column that holds it.  This is the
 * loop nest of RunPairWiseAlignments (include/overlap.hpp:518-546) without the host in between. */
int bella_xdrop_align_csc_device(bella_xdrop* h, uint32_t n_cols, const uint32_t* d_colptrC, uint64_t n_pairs,
                                 const uint32_t* d_rowids, const uint16_t* d_posH, const uint16_t* d_posV, int32_t* d_out);

/* stats[0] = milliseconds of the last batch's kernels (CUDA events), [1] = extensions that went through the wide path,
 * [2] = kernel launches of the last batch, [3] = lanes, [4] = cells per lane actually used. */
int bella_xdrop_get_stats(bella_xdrop* h, double* stats5);

void* bella_xdrop_stream(bella_xdrop* h);
int   bella_xdrop_sync(bella_xdrop* h);

#ifdef __cplusplus
}
#endif
#endif
