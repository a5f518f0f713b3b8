/* bella_b200.h -- C-ABI of the B200-native overlap-detection SpGEMM  C = A * A^T  for BELLA.
 *
 * Drop-in boundary for the reference's  estimateFLOP / estimateNNZ_Hash / LocalSpGEMM  trio inside
 * HashSpGEMM (reference include/overlap.hpp:157-363, called from include/overlap.hpp:667-719, whose
 * only call site is src/main.cpp:499-525).  The reference has no FFI of its own: the
 * binding a maintainer adds is the header-only shim bella_b200/csrc/overlap_b200.hpp, which has
 * HashSpGEMM's template signature, packs the CSC<uint32_t, unsigned short> members into
 * bella_csc_view, and calls the functions below.
 *
 * Conventions: plain C, POD only, `int` status return (0 = ok, negative = error, never exit()),
 * the caller owns every buffer it passes, one caller thread per handle, one handle per GPU.
 * There is no CPU fallback: every entry point fails with BELLA_B200_ERR_CUDA when no sm_100 device
 * is usable.
 *
 * Semantics (must match the reference bit-for-bit on the same arrays):
 *   - A = reads x k-mers (the reference's `spmat`,  src/main.cpp:489), CSC, rows unsorted in a column
 *   - B = k-mers x reads (the reference's `transpmat`, src/main.cpp:476), CSC; the ORDER of the
 *     nonzeros inside a column of B is the fold order of the semiring and is significant
 *   - values are k-mer start positions in the read (unsigned short)
 *   - only the strictly lower triangle is produced (row > col; include/overlap.hpp:315)
 *   - multiply = multiop / overlapop (include/chain.hpp:47-86), add = chainop
 *     (include/chain.hpp:100-150), final pick = spmatType_::choose() (include/common/common.h:162-170)
 *   - the substring comparison of checkstrand (include/chain.hpp:35-44) is replaced by per-nonzero
 *     strand bits: bit = 1 iff the k-mer window in the read equals the k-mer's canonical
 *     representative; oriented(product) = (strand_A[k] == strand_B[j]).
 *   - inside one output column the rows are returned in ascending order (the reference's order is
 *     hash-slot order and already schedule dependent).
 */
#ifndef BELLA_B200_H_
#define BELLA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BELLA_B200_OK            0
#define BELLA_B200_ERR_ARG      -1   /* bad argument / call order                                  */
#define BELLA_B200_ERR_CUDA     -2   /* CUDA runtime error or no usable device (no CPU fallback)     */
#define BELLA_B200_ERR_OOM      -3   /* device allocation failed                                     */
#define BELLA_B200_ERR_RANGE    -4   /* a count does not fit the reference's index types (u32 / u15) */
#define BELLA_B200_ERR_INTERNAL -5   /* inconsistency detected on the device                         */
#define BELLA_B200_ERR_CAPACITY -6   /* multi-GPU only: an exchange buffer sized by the caller is too small (every rank reports it) */

typedef struct bella_b200_handle bella_b200_handle;

/* Mirror of the public members of CSC<uint32_t, unsigned short> (reference include/common/CSC.h:88-95). */
typedef struct bella_csc_view {
	uint32_t rows, cols, nnz;
	const uint32_t* colptr;   /* [cols+1] */
	const uint32_t* rowids;   /* [nnz]    */
	const uint16_t* values;   /* [nnz]    */
} bella_csc_view;

/* Mirror of the public members of CSR<uint32_t, unsigned short> (reference include/common/CSR.h:15-67; the CSC -> CSR
 * conversion is src/CSR.cpp:59-91).  BELLA itself never instantiates CSR on this path, but a caller
 * that holds its read x k-mer matrix row-major can hand it over as is: the CSR arrays of A ARE the CSC arrays of
 * B = A^T (rowptr == B.colptr, colids == B.rowids), see bella_csr_as_transposed_csc below.  The order of the entries
 * inside a row is then the fold order of the semiring (what B's column order is for the CSC surface): the reference's
 * own conversion CSR(const CSC&) emits a row's entries by ascending column id, which is NOT the order MergeDuplicates
 * left in `transpmat`, so results equal the reference's only if the rows keep transpmat's column order. */
typedef struct bella_csr_view {
	uint32_t rows, cols, nnz;
	const uint32_t* rowptr;   /* [rows+1] */
	const uint32_t* colids;   /* [nnz]    */
	const uint16_t* values;   /* [nnz]    */
	int zerobased;            /* CSR::zerobased; one-based views (after ConvertOneBased(), CSR.h:41-49) are refused */
} bella_csr_view;

/* Zero-copy reinterpretation: CSR of A (reads x k-mers) -> CSC of B = A^T (k-mers x reads).  Returns 0, or
 * BELLA_B200_ERR_ARG for a one-based view. */
static inline int bella_csr_as_transposed_csc(const bella_csr_view* A_csr, bella_csc_view* B_out)
{
	if (!A_csr || !B_out || !A_csr->zerobased) return -1;
	B_out->rows = A_csr->cols; B_out->cols = A_csr->rows; B_out->nnz = A_csr->nnz;
	B_out->colptr = A_csr->rowptr; B_out->rowids = A_csr->colids; B_out->values = A_csr->values;
	return 0;
}

/* One handle per GPU.  device = CUDA ordinal. */
int bella_b200_create(bella_b200_handle** out, int device);
int bella_b200_destroy(bella_b200_handle* h);
const char* bella_b200_last_error(const bella_b200_handle* h);

/* Inputs from HOST memory.  The copy to the device is started here on the handle's own copy stream, in a few
 * ranges of reads, and bella_b200_symbolic consumes each range as soon as it has landed (the transpose overlaps
 * the upload): the host arrays must stay valid and unchanged until bella_b200_symbolic has returned.
 * Page-locked host memory makes the copies truly asynchronous; pageable memory works, without the overlap.  Replaces the `A, B, reads, bpars` arguments
 * of HashSpGEMM (include/overlap.hpp:650-652): read_len[n] and the strand bits stand in for `reads`,
 * kmer_size/bin_size for BELLApars.{kmerSize,binSize}.  strand bits are bit-packed, LSB first, one
 * bit per nonzero in that matrix's array order.  strand_B may be NULL: then bit 31 of every
 * B.rowids entry is that nonzero's strand bit and the k-mer id is the low 31 bits (the compressed
 * panel format the multi-GPU all-gather exchanges; needs fewer than 2^31 k-mers).
 * A == B^T is what BELLA always passes (src/main.cpp:489), so the device derives A from B itself
 * (it needs each nonzero's position inside B's column, which A's arrays do not carry): A may be
 * NULL; when given, only its shape is checked against B and its arrays and strand_A are never
 * read or copied.  B.cols == number of reads n, B.rows == number of k-mers m. */
int bella_b200_set_inputs(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B,
		const uint32_t* read_len, const uint8_t* strand_A, const uint8_t* strand_B,
		uint16_t kmer_size, uint16_t bin_size);

/* The CSR surface of the same call: A given row-major (reads x k-mers, HOST arrays).  Equivalent to
 * bella_b200_set_inputs(h, NULL, B, ...) with B = bella_csr_as_transposed_csc(A_csr); strand bits are per nonzero in the
 * CSR array order. */
int bella_b200_set_inputs_csr(bella_b200_handle* h, const bella_csr_view* A_csr, const uint32_t* read_len,
		const uint8_t* strand, uint16_t kmer_size, uint16_t bin_size);

/* Same, but every pointer (including those inside the views) is a DEVICE pointer on the handle's
 * GPU; nothing is copied and the caller keeps the buffers alive until the next set_inputs/destroy.
 * This is the entry the multi-GPU path uses after the NCCL all-gather of the B panel. */
int bella_b200_set_inputs_device(bella_b200_handle* h, const bella_csc_view* A, const bella_csc_view* B,
		const uint32_t* read_len, const uint8_t* strand_A, const uint8_t* strand_B,
		uint16_t kmer_size, uint16_t bin_size);

/* Restrict this handle to output columns [col_lo, col_hi) (row sharding across GPUs).
 * Default after set_inputs: [0, n). */
int bella_b200_set_column_range(bella_b200_handle* h, uint32_t col_lo, uint32_t col_hi);

/* Symbolic phase == estimateFLOP + prefixsum + estimateNNZ_Hash + prefixsum
 * (include/overlap.hpp:667-679) for the handle's column range.
 *   flops    (nullable) total number of kept products (64-bit; the reference's is uint32_t)
 *   flopC    (nullable) HOST [col_hi-col_lo] per-column product count   == estimateFLOP()
 *   colptrC  (nullable) HOST [col_hi-col_lo+1] exclusive scan of per-column nnz(C), colptrC[0]=0
 *                       == prefixsum(estimateNNZ_Hash()) restricted to the range */
int bella_b200_symbolic(bella_b200_handle* h, uint64_t* flops, uint32_t* flopC, uint32_t* colptrC);

/* Numeric phase == LocalSpGEMM(col_begin, col_end, ...) + choose() (include/overlap.hpp:281-363,
 * include/common/common.h:162-170).  (The device computes the values together with the structure
 * during bella_b200_symbolic; this call compacts them into C's arrays and copies the range out.)  [col_begin, col_end) are GLOBAL column ids inside the handle's
 * range.  Outputs are HOST arrays of colptrC[col_end]-colptrC[col_begin] entries; column i's
 * entries start at colptrC[i]-colptrC[col_begin], rows ascending.
 *   rowidsC : row read id (the "H" read, larger id)
 *   count   : spmatType_::count (u16 recurrence of chain.hpp:105,140)
 *   posH/posV : first k-mer of the most supported bin (position on the row / column read)
 * Any output pointer may be NULL. */
int bella_b200_numeric(bella_b200_handle* h, uint32_t col_begin, uint32_t col_end,
		uint32_t* rowidsC, uint16_t* count, uint16_t* posH, uint16_t* posV);

/* Optional, before bella_b200_symbolic: HOST output buffers of `capacity` entries each (page-locked for a truly asynchronous
 * copy) that the caller will pass to bella_b200_numeric for the handle's WHOLE column range.  The symbolic phase then compacts
 * and copies the results of a range of columns as soon as that range is folded, while the later ranges still fold, and the
 * bella_b200_numeric call with these same pointers has nothing left to copy.  The reference sizes its outputs after the symbolic
 * phase (include/overlap.hpp:716-741); a caller that knows a bound of nnz(C) -- the previous batch's, or the product count --
 * saves the exposed device->host copy.  A result larger than `capacity` is NOT written: bella_b200_numeric then copies as
 * usual into whatever it is given.  rowidsC == NULL switches it off. */
int bella_b200_set_output_buffers(bella_b200_handle* h, uint32_t* rowidsC, uint16_t* count, uint16_t* posH, uint16_t* posV, uint64_t capacity);

/* Optional per-nonzero extras for the same range (HOST, nullable each):
 *   nbins    number of bins of the pair's final value
 *   support  support of the chosen bin            (spmatType_::chain(), common.h:142-149)
 *   overlap  overlap estimate of the chosen bin   (spmatType_::overlaplength(), common.h:152-159) */
int bella_b200_numeric_aux(bella_b200_handle* h, uint32_t col_begin, uint32_t col_end,
		uint16_t* nbins, uint16_t* support, uint16_t* overlap);

/* Kept products of the handle's column range (== sum of estimateFLOP's array over the range), after the symbolic phase
 * or bella_b200_mg_finish. */
int bella_b200_get_flops(bella_b200_handle* h, uint64_t* flops);

/* Number of output nonzeros of the handle's column range whose final value has MORE THAN 16 bins.  For those the
 * reference's choose() (include/common/common.h:162-170) depends on the tie order of libstdc++'s introsort, which is
 * pinned (insertion sort, stable) only up to 16 elements: the device picks the best-supported bin with ties to the most
 * recent one, exactly as for fewer bins, and reports here how many pairs that rule was not validated for (0 on every
 * input seen so far; the CPU oracle counts the same thing).  Valid after the numeric phase. */
int bella_b200_n_unpinned(bella_b200_handle* h, uint64_t* n_unpinned);

/* Device-resident results of the whole column range after bella_b200_numeric_device():
 * runs the numeric phase without any device->host copy. */
int bella_b200_numeric_device(bella_b200_handle* h);
int bella_b200_result_device(bella_b200_handle* h, const uint32_t** colptrC, const uint32_t** rowidsC,
		const uint16_t** count, const uint16_t** posH, const uint16_t** posV, uint64_t* nnzC);

/* One full pass (layout + symbolic + numeric) over device-resident inputs with no host
 * synchronisation except the final one; used by bench.py for the device-resident figure.
 * nnzC_out / flops_out nullable. */
int bella_b200_run_resident(bella_b200_handle* h, uint64_t* nnzC_out, uint64_t* flops_out);

/* Timings of the last pass in milliseconds (CUDA events on the handle's streams):
 *   [0] two-level partition (k_rp1 + k_rp2)   [6] k_bucket + plan kernels
 *   [1] the scatter | group + fold pipeline (the scatter passes of the later column ranges run beside the group + fold
 *       kernels of the earlier ones; all capacity classes)   [7] the scatter passes alone, first to last (inside [1])
 *   [2] output (C's colptr scans + k_compact)   [3] host->device copies   [4] device->host copies   [5] kernels launched (count) */
int bella_b200_get_timings(bella_b200_handle* h, float* ms8);

/* cudaStream_t of the handle as an opaque pointer (so a caller can order its own work after it). */
void* bella_b200_stream(bella_b200_handle* h);
/* Run all further work of the handle on the caller's stream (a cudaStream_t; NULL = the legacy default
 * stream).  The handle does not take ownership. */
int bella_b200_set_stream(bella_b200_handle* h, void* stream);

/* ---- multi-GPU, one handle (process) per GPU; the caller runs the collectives between the calls.
 * After the all-gather of the B panel every GPU holds the whole B (bella_b200_set_inputs_device).  It then
 *   mg_transpose : transposes only the k-mers [kmer_lo, kmer_hi) (all reads) and writes, per output column of
 *                  the WHOLE matrix, how many kept products those k-mers contribute: cnt_local_dev u32[n] (device)
 *   mg_scatter   : expands those products into sendbuf_dev (device, 8 bytes each) ordered by output column;
 *                  sendoff_dev u64[n+1] (device) is the exclusive scan of cnt_local_dev
 *   -- the caller all-gathers the counts and exchanges the products (all-to-all) so that the owner of an
 *      output column range receives, per source GPU, that range's products in column order --
 *   mg_finish    : for the columns [col_lo, col_hi) this GPU owns: counts_all_dev u32[world][n] (every GPU's
 *                  cnt_local), recv_dev = the received products, source-major; recvbase_dev u64[world] = offset of
 *                  each source's block in recv_dev; segoff_dev u64[world][ncols+1] = per source the exclusive scan
 *                  of its counts over the owned columns.  Runs the plan, group + fold and leaves the handle in
 *                  the state bella_b200_symbolic leaves it in (bella_b200_numeric / _numeric_device / _get_colptr).
 * All pointers are device pointers on the handle's GPU and stay valid during the call. */
int bella_b200_mg_transpose(bella_b200_handle* h, uint32_t kmer_lo, uint32_t kmer_hi, uint32_t* cnt_local_dev);
int bella_b200_mg_scatter(bella_b200_handle* h, const uint64_t* sendoff_dev, uint64_t* sendbuf_dev);
int bella_b200_mg_finish(bella_b200_handle* h, uint32_t col_lo, uint32_t col_hi, int world, const uint32_t* counts_all_dev,
		const uint64_t* segoff_dev, const uint64_t* recvbase_dev, const uint64_t* recv_dev);
/* NVLink mode (bella_b200/distributed.py mode "nvlink"): no collective on the data path.  Every rank maps the others'
 * exchange buffers (CUDA peer memory over NVLink / NVSwitch; the Python side allocates them as torch symmetric memory)
 * and the kernels store straight into the owner's memory; between the phases the caller only runs a barrier.
 *   mg_geometry         the two-level partition all ranks agree on: out8 = {wshift, shift1, nb1 (coarse buckets), nb1_loc
 *                       (coarse buckets per rank), k-mers per rank, cap1 (records per sub-region), fine buckets, l2}
 *   mg_route_push       level 1 of the transpose of this rank's reads [read_lo, read_hi) (colptr indexed by the GLOBAL read
 *                       id, strand bit in bit 31 of the row ids): every nonzero goes into the coarse bucket of its k-mer on
 *                       the rank that transposes that k-mer range -- sub-region `me` of the bucket, in peer_E/peer_K[rank] --
 *                       and the record counts into peer_cnt[rank]                       -- barrier --
 *   mg_transpose_coarse level 2 + bucket kernel over the received coarse buckets; cnt_local_dev u32[n] as mg_transpose
 *   mg_post             a 32-bit array of this rank to the same offset `at` on every rank (remote stores)
 *   mg_exchange         phase 0: this rank's per-column counts -> row `me` of counts_all on every rank   -- barrier --
 *                       (counts_all is u32 [world * n] + 16 words + one error word per rank; push_dev is u64 [3 * world + 2]: the last
 *                       two words receive the largest receive / send block any rank needs)
 *                       phase 1: exchange plan on the device (flat scan of counts_all + one kernel: send offsets, per-source
 *                       segment offsets, receive bases, push blocks), expansion of this rank's products into its send
 *                       buffer (mg_scatter), push of each destination's block into its receive buffer   -- barrier --
 *   then mg_finish on the receive buffer. */
int bella_b200_mg_geometry(uint32_t n_kmers, uint64_t nnz_total, int world, uint32_t* out8);
int bella_b200_mg_route_push(bella_b200_handle* h, uint32_t read_lo, uint32_t read_hi, const uint32_t* colptr_global_dev, const uint32_t* rowids_dev,
		const uint16_t* values_dev, uint32_t n_kmers, const uint32_t* geom8, int world, int me, void* const* peer_E, void* const* peer_K, void* const* peer_cnt);
int bella_b200_mg_transpose_coarse(bella_b200_handle* h, uint32_t kmer_lo, uint32_t kmer_hi, const uint32_t* geom8, int world,
		const uint64_t* E_dev, const uint32_t* K_dev, const uint32_t* cnt_dev, uint64_t nnz_cap, uint32_t* cnt_local_dev);
int bella_b200_mg_post(bella_b200_handle* h, uint64_t count, const uint32_t* src_dev, int world, uint64_t at, void* const* peer_dst);
int bella_b200_mg_exchange(bella_b200_handle* h, int world, int me, const uint32_t* cuts_dev, const uint32_t* cnt_local_dev, void* const* peer_counts_all,
		int phase, const uint32_t* counts_all_dev, uint64_t* scan_dev, uint64_t cap_recv, uint64_t cap_send, uint64_t* sendoff_dev, uint64_t* segoff_dev,
		uint64_t* recvbase_dev, uint64_t* push_dev, uint64_t* sendbuf_dev, void* const* peer_recv);
/* colptrC of the handle's column range to HOST memory ([col_hi-col_lo+1]) once the symbolic phase has run. */
int bella_b200_get_colptr(bella_b200_handle* h, uint32_t* colptrC_host);

/* ---- "next" row f2: matrix construction on the device.
 * Replaces  CSC<IT,NT> transpmat(alltuples, nkmer, numReads, keep-p1, needsort=false)  (src/main.cpp:476-480 ->
 * src/CSC.cpp:422-479: counting sort by read, then MergeDuplicates :301-420) and  transpmat.Transpose()
 * (src/main.cpp:489).  Tuples are (k-mer id, read id, position) in HOST memory, the tuples of one read contiguous and in
 * position order (what src/main.cpp:393-416 emits; anything else is refused), t_strand one bit per tuple (LSB first).
 * B comes out in exactly the reference's order -- the hash-slot order of MergeDuplicates, which is the fold order of
 * the SpGEMM -- stays on the device, and the handle is left as after bella_b200_set_inputs. */
int bella_b200_set_inputs_tuples(bella_b200_handle* h, uint32_t n_kmers, uint32_t n_reads, uint64_t ntuples, const uint32_t* t_kmer,
		const uint32_t* t_read, const uint16_t* t_pos, const uint8_t* t_strand, const uint32_t* read_len, uint16_t kmer_size, uint16_t bin_size);
/* The handle's B back to HOST memory (any pointer may be NULL; call with NULLs first for nnz): colptr [n+1], rowids/values
 * [nnz], strand bits [(nnz+7)/8]; build_ms = device time of the last bella_b200_set_inputs_tuples without its upload. */
int bella_b200_get_B(bella_b200_handle* h, uint32_t* nnz, uint32_t* colptr_host, uint32_t* rowids_host, uint16_t* values_host,
		uint8_t* strand_host, float* build_ms);

#ifdef __cplusplus
}
#endif
#endif /* BELLA_B200_H_ */
