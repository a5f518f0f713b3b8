/* bella_kmers.h -- C-ABI of the "next" row f3: reliable k-mer selection and tuple emission on one B200.
 *
 * Stands for, in the reference:
 *   SplitCount(allfiles, countsreliable, lower, upper, upperlimit, bpars)      include/kmercount.hpp:466-677
 *        (called at src/main.cpp:299-302; HyperLogLog -> Bloom filter -> cuckoo-hash counts -> l <= count <= u)
 *   the tuple emission loop                                                    src/main.cpp:339-423
 *        (per read, per position: canonical k-mer -> dictionary lookup -> (kmer_id, read_id, pos))
 * The result is what that loop leaves in `alltuples`: one (k-mer id, read, pos) per occurrence of a reliable k-mer, the tuples
 * of a read contiguous and in position order, plus the strand bit the overlap SpGEMM needs (1 iff the window equals its
 * canonical representative).  K-mer ids are cuckoo iteration order in the reference, i.e. arbitrary; here id = rank of the
 * canonical k-mer's packed value.  The id-free content -- which (read, pos) are emitted, how many distinct k-mers -- is
 * identical to the reference's.  Output feeds bella_b200_set_inputs_tuples (include/bella_b200.h, row f2) unchanged.
 *
 * Plain pointers and sizes; 0 or a negative BELLA_KMERS_E* code; bella_kmers_last_error() explains.  No CPU fallback.
 * Implemented by bella_b200/libbella_kmers.so (bella_b200/csrc/bella_kmers.cu + kmers.cuh).
 * STATUS: written after round 1's GPU time was spent -- logic checked on the CPU (tests/test_kmers_host.py runs the same
 * per-element functions), not yet run or measured on a B200. */
#ifndef BELLA_KMERS_H
#define BELLA_KMERS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BELLA_KMERS_OK       0
#define BELLA_KMERS_EINVAL  (-1)
#define BELLA_KMERS_ECUDA   (-2)
#define BELLA_KMERS_ERANGE  (-3)   /* more than 2^31 - 1 bases, or more than 2^32 - 2 reliable k-mers */

typedef struct bella_kmers bella_kmers;

bella_kmers* bella_kmers_create(int device);
void         bella_kmers_destroy(bella_kmers* h);
const char*  bella_kmers_last_error(const bella_kmers* h);

/* reads: concatenated, one byte per base as they come from the FASTQ file, read i = seqs[seq_off[i] .. seq_off[i+1]);
 * k = BELLApars.kmerSize (<= 32), [lower, upper] = the reliable range (src/main.cpp:256-276).  Results stay on the device. */
int bella_kmers_count(bella_kmers* h, const char* seqs, const uint64_t* seq_off, uint32_t n_reads, int k, int lower, int upper,
                      uint64_t* n_kmers, uint64_t* n_tuples);

/* the tuples to host arrays of n_tuples entries; t_strand = one bit per tuple, LSB first ((n_tuples + 7) / 8 bytes), nullable */
int bella_kmers_get_tuples(bella_kmers* h, uint32_t* t_kmer, uint32_t* t_read, uint16_t* t_pos, uint8_t* t_strand_bits);

/* stats[0] = milliseconds of the last count's kernels (CUDA events), [1] = positions keyed, [2] = kernel launches */
int bella_kmers_get_stats(bella_kmers* h, double* stats3);

#ifdef __cplusplus
}
#endif
#endif
