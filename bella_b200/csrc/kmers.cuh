// "Next" row f3: reliable k-mer selection and tuple emission on the device.
//
// What it replaces: SplitCount (include/kmercount.hpp:466-677 -- HyperLogLog estimate, Bloom filter, cuckoo-hash counts,
// keep l <= count <= u) and the tuple emission loop of src/main.cpp:339-423.  The HLL / Bloom / cuckoo machinery is a memory
// optimisation of "count every canonical k-mer exactly"; with 180 GB of HBM the exact statement is cheaper: one key per
// read position, one sort, and element-wise passes over the sorted array.
//
//   extract   key[g] = canonical k-mer starting at base g of the concatenated reads (2 bits per base as Kmer::set_kmer packs
//             them, kmercode/Kmer.cpp:215-216; the smaller of the k-mer and its reverse complement, Kmer::rep()), or the
//             sentinel ~0 where the window would cross a read end; val[g] = g | strand << 31
//   sort      (key, val) pairs by key                                   [cub radix sort: plumbing]
//   classify  an element is reliable iff its run of equal keys has l..u members: the run head counts forward at most u
//             members, every other member looks back at most u-1 places for its head
//   ids       exclusive scan over the reliable run heads                [cub scan]: id = rank of the canonical k-mer
//   place     id_at[g] = id of the reliable k-mer at position g, NONE elsewhere
//   emit      scan over id_at != NONE in position order [cub scan], then (k-mer id, read, pos, strand) per occurrence:
//             the tuples of a read contiguous and in position order, which is what src/main.cpp:393-416 emits and what
//             bella_b200_set_inputs_tuples (row f2) takes.
//
// Every kernel is one independent function per element (KM_FN below), so this header also compiles under g++: the test
// harness tests/emu/kmers_host.cpp runs the same functions in plain loops (std::sort / prefix sums in place of cub)
// against the CPU restatement oracle_reliable_occurrences.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KM_FN __host__ __device__ __forceinline__
#else
#define KM_FN inline
#endif

namespace km {

constexpr uint64_t SENTINEL = ~0ull;           // no k-mer starts here (k <= 31, or k = 32 and never all-T on both strands: see valid())
constexpr uint32_t NONE = 0xffffffffu;

// Kmer::set_kmer's base code (kmercode/Kmer.cpp:215-216): A 0, C 1, G 2, T 3; lower case alike; any other byte by the same two
// bits (N counts as G) -- the reference has no N check on this path
KM_FN uint64_t base_code(char c) { const unsigned x = ((unsigned)c & 4u) >> 1; return x + ((x ^ ((unsigned)c & 2u)) >> 1); }

// read that holds global base index g: the last r with seq_off[r] <= g
KM_FN uint32_t read_of(const uint64_t* seq_off, uint32_t n_reads, uint64_t g)
{
	uint32_t lo = 0, hi = n_reads;               // seq_off[lo] <= g < seq_off[hi]
	while (hi - lo > 1) {
		const uint32_t mid = lo + ((hi - lo) >> 1);
		if (seq_off[mid] <= g) lo = mid; else hi = mid;
	}
	return lo;
}

// one element of `extract`
KM_FN void extract_one(uint64_t g, const char* seqs, const uint64_t* seq_off, uint32_t n_reads, int k, uint64_t* key, uint32_t* val)
{
	const uint32_t r = read_of(seq_off, n_reads, g);
	if (g + (uint64_t)k > seq_off[r + 1]) { key[g] = SENTINEL; val[g] = (uint32_t)g; return; }
	uint64_t fw = 0, rv = 0;
	for (int i = 0; i < k; ++i) {
		const uint64_t c = base_code(seqs[g + i]);
		fw = (fw << 2) | c;
		rv |= (3 - c) << (2 * i);
	}
	const bool fwd_is_canon = fw <= rv;
	key[g] = fwd_is_canon ? fw : rv;
	val[g] = (uint32_t)g | (fwd_is_canon ? 0x80000000u : 0u);
}

// one element of `classify` on the sorted keys: head[i] = 1 iff i starts a reliable run, rel[i] = 1 iff i belongs to one
KM_FN void classify_one(uint64_t i, uint64_t n, const uint64_t* key, int lower, int upper, uint8_t* head, uint8_t* rel)
{
	const uint64_t me = key[i];
	head[i] = 0; rel[i] = 0;
	if (me == SENTINEL) return;
	uint64_t h = i;                              // look back for the run head, at most upper - 1 places
	int back = 0;
	while (h > 0 && key[h - 1] == me) {
		if (++back >= upper) return;             // at least upper + 1 members: not reliable
		--h;
	}
	int count = back + 1;                        // members h..i; count on to the right, stopping once the run is too long
	for (uint64_t j = i + 1; j < n && key[j] == me; ++j)
		if (++count > upper) return;
	if (count < lower) return;
	rel[i] = 1;
	if (h == i) head[i] = 1;
}

// one element of `place`: scan[i] = number of reliable heads before i (exclusive), so a member's id is the inclusive count - 1
KM_FN void place_one(uint64_t i, const uint32_t* val, const uint8_t* head, const uint8_t* rel, const uint32_t* scan, uint32_t* id_at)
{
	const uint32_t g = val[i] & 0x7fffffffu;
	id_at[g] = rel[i] ? scan[i] + head[i] - 1 : NONE;
}

// `place` needs the strand of position g later; it travels in a second array written here
KM_FN void strand_one(uint64_t i, const uint32_t* val, uint8_t* strand_at) { strand_at[val[i] & 0x7fffffffu] = (uint8_t)(val[i] >> 31); }

// one element of `emit`: slot[g] = number of reliable positions before g (exclusive scan of id_at != NONE)
KM_FN void emit_one(uint64_t g, const uint32_t* id_at, const uint8_t* strand_at, const uint64_t* slot, const uint64_t* seq_off, uint32_t n_reads,
		uint32_t* t_kmer, uint32_t* t_read, uint16_t* t_pos, uint8_t* t_strand)
{
	if (id_at[g] == NONE) return;
	const uint64_t t = slot[g];
	const uint32_t r = read_of(seq_off, n_reads, g);
	t_kmer[t] = id_at[g]; t_read[t] = r; t_pos[t] = (uint16_t)(g - seq_off[r]); t_strand[t] = strand_at[g];
}

}  // namespace km
