// TEST INFRASTRUCTURE.  Host run of bella_b200/csrc/kmers.cuh: the per-element functions the f3 kernels are made of, executed
// in plain loops with std::sort and prefix sums where the device uses cub.  Same stages, same arrays as bella_kmers.cu's
// bella_kmers_count / _get_tuples, so the logic is checked against the oracle without a GPU.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "../../bella_b200/csrc/kmers.cuh"

extern "C" int kmers_host_count(const char* seqs, const uint64_t* seq_off, uint32_t n_reads, int k, int lower, int upper,
		uint32_t* t_kmer, uint32_t* t_read, uint16_t* t_pos, uint8_t* t_strand_bits, uint64_t cap, uint64_t* n_kmers, uint64_t* n_tuples)
{
	const uint64_t n = seq_off[n_reads];
	*n_kmers = *n_tuples = 0;
	if (n == 0) return 0;
	std::vector<uint64_t> key(n); std::vector<uint32_t> val(n);
	for (uint64_t g = 0; g < n; ++g) km::extract_one(g, seqs, seq_off, n_reads, k, key.data(), val.data());
	std::vector<uint64_t> order(n);
	std::iota(order.begin(), order.end(), 0);
	std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t b) { return key[a] < key[b]; });      // not stable, like the device sort's ties
	std::vector<uint64_t> skey(n); std::vector<uint32_t> sval(n);
	for (uint64_t i = 0; i < n; ++i) { skey[i] = key[order[i]]; sval[i] = val[order[i]]; }
	std::vector<uint8_t> head(n), rel(n);
	for (uint64_t i = 0; i < n; ++i) km::classify_one(i, n, skey.data(), lower, upper, head.data(), rel.data());
	std::vector<uint32_t> scan(n);
	uint32_t acc = 0;
	for (uint64_t i = 0; i < n; ++i) { scan[i] = acc; acc += head[i]; }
	*n_kmers = acc;
	std::vector<uint32_t> id_at(n); std::vector<uint8_t> strand_at(n);
	for (uint64_t i = 0; i < n; ++i) { km::place_one(i, sval.data(), head.data(), rel.data(), scan.data(), id_at.data()); km::strand_one(i, sval.data(), strand_at.data()); }
	std::vector<uint64_t> slot(n);
	uint64_t nt = 0;
	for (uint64_t g = 0; g < n; ++g) { slot[g] = nt; nt += id_at[g] != km::NONE; }
	*n_tuples = nt;
	if (nt > cap) return -1;
	std::vector<uint8_t> strand(nt ? nt : 1);
	for (uint64_t g = 0; g < n; ++g) km::emit_one(g, id_at.data(), strand_at.data(), slot.data(), seq_off, n_reads, t_kmer, t_read, t_pos, strand.data());
	for (uint64_t b = 0; b < (nt + 7) / 8; ++b) {
		unsigned v = 0;
		for (int j = 0; j < 8; ++j) { const uint64_t t = b * 8 + j; if (t < nt && strand[t]) v |= 1u << j; }
		t_strand_bits[b] = (uint8_t)v;
	}
	return 0;
}
