// Host-side planning of one pairwise contraction: label classification -> M/N/K/batch mode groups ->
// offset tables (or affine strides) for the permutation-fused GEMM.
#pragma once
#include <vector>

#include "gemm_c128.cuh"

namespace qb {

struct ModeGroup {
    // merged modes of one class, fastest first; strides per operand (0 = operand does not hold the mode)
    std::vector<int64_t> ext, sa, sb, sc;
    int64_t total() const {
        int64_t t = 1;
        for (auto e : ext) t *= e;
        return t;
    }
};

struct ContractSpec {
    ModeGroup m, n, k, b;
    bool swapped = false;  // A and B roles exchanged so that the larger free dimension sits on the 128-wide side
    int conjA = 0, conjB = 0;
};

// returns QB200_OK or an error code with message in *err
int32_t make_contract_spec(int rankA, const int64_t* extA, const int32_t* modesA, int rankB, const int64_t* extB,
                           const int32_t* modesB, int rankC, const int64_t* extC, const int32_t* modesC,
                           ContractSpec* spec, std::string* err);

// number of int64 table entries the spec needs on the device
int64_t contract_table_entries(const ContractSpec& s);

// fills GemmArgs (except data pointers / alpha / beta) and launches the table builders into `tables`
int32_t materialize_contract(qb200_ctx* ctx, const ContractSpec& s, int64_t* tables, GemmArgs* g);

}  // namespace qb
