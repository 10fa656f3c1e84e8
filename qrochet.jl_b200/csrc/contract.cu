// qb200_contract: Tenet.contract(a, b; dims) on device tensors (call sites: Chain.jl:372,602,616,636,682,
// 734,747; every node of the EinExprs path in examples/distributed.jl:89).
#include "contract.cuh"

#include <algorithm>

namespace qb {

static int find_mode(const int32_t* modes, int rank, int32_t label) {
    for (int i = 0; i < rank; ++i)
        if (modes[i] == label) return i;
    return -1;
}

static void dense_strides(int rank, const int64_t* ext, int64_t* st) {
    int64_t s = 1;
    for (int i = 0; i < rank; ++i) {
        st[i] = s;
        s *= ext[i];
    }
}

static void push_mode(ModeGroup& g, int64_t ext, int64_t sa, int64_t sb, int64_t sc) {
    if (ext == 1) return;  // size-1 modes carry no offset
    if (!g.ext.empty()) {
        // merge with the previous mode when it is contiguous after it in every operand that holds it
        size_t j = g.ext.size() - 1;
        int64_t e = g.ext[j];
        if (sa == g.sa[j] * e && sb == g.sb[j] * e && sc == g.sc[j] * e) {
            g.ext[j] *= ext;
            return;
        }
    }
    g.ext.push_back(ext);
    g.sa.push_back(sa);
    g.sb.push_back(sb);
    g.sc.push_back(sc);
}

int32_t make_contract_spec(int rankA, const int64_t* extA, const int32_t* modesA, int rankB, const int64_t* extB,
                           const int32_t* modesB, int rankC, const int64_t* extC, const int32_t* modesC,
                           ContractSpec* spec, std::string* err) {
    auto fail = [&](const char* msg) {
        if (err) *err = msg;
        return (int32_t)QB200_E_INVALID;
    };
    for (int i = 0; i < rankA; ++i)
        for (int j = i + 1; j < rankA; ++j)
            if (modesA[i] == modesA[j]) return fail("contract: repeated mode in A");
    for (int i = 0; i < rankB; ++i)
        for (int j = i + 1; j < rankB; ++j)
            if (modesB[i] == modesB[j]) return fail("contract: repeated mode in B");
    for (int i = 0; i < rankC; ++i)
        for (int j = i + 1; j < rankC; ++j)
            if (modesC[i] == modesC[j]) return fail("contract: repeated mode in C");

    int64_t stA[QB200_MAX_RANK], stB[QB200_MAX_RANK], stC[QB200_MAX_RANK];
    dense_strides(rankA, extA, stA);
    dense_strides(rankB, extB, stB);
    dense_strides(rankC, extC, stC);

    // free sizes decide the orientation
    int64_t Mtot = 1, Ntot = 1;
    for (int i = 0; i < rankA; ++i)
        if (find_mode(modesB, rankB, modesA[i]) < 0 && find_mode(modesC, rankC, modesA[i]) >= 0) Mtot *= extA[i];
    for (int i = 0; i < rankB; ++i)
        if (find_mode(modesA, rankA, modesB[i]) < 0 && find_mode(modesC, rankC, modesB[i]) >= 0) Ntot *= extB[i];
    bool swap = Ntot > Mtot;
    if (swap) {
        std::swap(rankA, rankB);
        std::swap(extA, extB);
        std::swap(modesA, modesB);
        for (int i = 0; i < QB200_MAX_RANK; ++i) std::swap(stA[i], stB[i]);
    }
    ContractSpec s;
    s.swapped = swap;

    for (int i = 0; i < rankC; ++i)
        if (find_mode(modesA, rankA, modesC[i]) < 0 && find_mode(modesB, rankB, modesC[i]) < 0)
            return fail("contract: output mode not present in any input");

    // M modes in A's order, K modes in A's order (then B-only summed modes), N in B's order, batch in C's order
    for (int i = 0; i < rankA; ++i) {
        int jb = find_mode(modesB, rankB, modesA[i]), jc = find_mode(modesC, rankC, modesA[i]);
        if (jb >= 0 && extB[jb] != extA[i]) return fail("contract: extent mismatch between A and B");
        if (jc >= 0 && extC[jc] != extA[i]) return fail("contract: extent mismatch between A and C");
        if (jb < 0 && jc >= 0) push_mode(s.m, extA[i], stA[i], 0, stC[jc]);
    }
    for (int i = 0; i < rankA; ++i) {
        int jb = find_mode(modesB, rankB, modesA[i]), jc = find_mode(modesC, rankC, modesA[i]);
        if (jc < 0) push_mode(s.k, extA[i], stA[i], jb >= 0 ? stB[jb] : 0, 0);
    }
    for (int i = 0; i < rankB; ++i) {
        int ja = find_mode(modesA, rankA, modesB[i]), jc = find_mode(modesC, rankC, modesB[i]);
        if (jc >= 0 && extC[jc] != extB[i]) return fail("contract: extent mismatch between B and C");
        if (ja < 0 && jc >= 0) push_mode(s.n, extB[i], 0, stB[i], stC[jc]);
        if (ja < 0 && jc < 0) push_mode(s.k, extB[i], 0, stB[i], 0);
    }
    for (int i = 0; i < rankC; ++i) {
        int ja = find_mode(modesA, rankA, modesC[i]), jb = find_mode(modesB, rankB, modesC[i]);
        if (ja >= 0 && jb >= 0) push_mode(s.b, extC[i], stA[ja], stB[jb], stC[i]);
    }
    for (ModeGroup* g : {&s.m, &s.n, &s.k, &s.b})
        if (g->ext.size() > 32) return fail("contract: more than 32 non-mergeable modes in one group");
    if (s.m.total() > INT32_MAX || s.n.total() > INT32_MAX || s.k.total() > INT32_MAX)
        return fail("contract: M/N/K extent exceeds int32");
    if (s.b.total() > 65535) return fail("contract: more than 65535 batch entries");
    *spec = s;
    return QB200_OK;
}

static int64_t entries_of(const ModeGroup& g, int noperands) { return g.ext.size() > 1 ? g.total() * noperands : 0; }

int64_t contract_table_entries(const ContractSpec& s) {
    return entries_of(s.m, 2) + entries_of(s.n, 2) + entries_of(s.k, 2) + entries_of(s.b, 3);
}

static int32_t make_operand(qb200_ctx* ctx, const ModeGroup& g, const std::vector<int64_t>& stride, int64_t*& cursor,
                            Operand* op) {
    if (g.ext.empty()) {
        *op = {nullptr, 0};
    } else if (g.ext.size() == 1) {
        *op = {nullptr, stride[0]};
    } else {
        ModeList ml;
        ml.n = (int)g.ext.size();
        for (int j = 0; j < ml.n; ++j) {
            ml.ext[j] = g.ext[j];
            ml.stride[j] = stride[j];
        }
        int64_t total = g.total();
        QB_TRY(build_offsets(ctx, ml, total, cursor));
        *op = {cursor, 0};
        cursor += total;
    }
    return QB200_OK;
}

int32_t materialize_contract(qb200_ctx* ctx, const ContractSpec& s, int64_t* tables, GemmArgs* g) {
    int64_t* cur = tables;
    QB_TRY(make_operand(ctx, s.m, s.m.sa, cur, &g->am));
    QB_TRY(make_operand(ctx, s.m, s.m.sc, cur, &g->cm));
    QB_TRY(make_operand(ctx, s.n, s.n.sb, cur, &g->bn));
    QB_TRY(make_operand(ctx, s.n, s.n.sc, cur, &g->cn));
    QB_TRY(make_operand(ctx, s.k, s.k.sa, cur, &g->ak));
    QB_TRY(make_operand(ctx, s.k, s.k.sb, cur, &g->bk));
    QB_TRY(make_operand(ctx, s.b, s.b.sa, cur, &g->ab));
    QB_TRY(make_operand(ctx, s.b, s.b.sb, cur, &g->bb));
    QB_TRY(make_operand(ctx, s.b, s.b.sc, cur, &g->cb));
    g->M = (int)s.m.total();
    g->N = (int)s.n.total();
    g->K = (int)s.k.total();
    g->batch = (int)s.b.total();
    // loader mapping: walk the dimension whose fastest mode is unit-stride in the operand
    g->a_kfast = (!s.k.ext.empty() && s.k.sa[0] == 1) ? 1 : 0;
    g->b_kfast = (!s.k.ext.empty() && s.k.sb[0] == 1) ? 1 : 0;
    return QB200_OK;
}

}  // namespace qb

extern "C" int32_t qb200_contract(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* modesA, int32_t conjA,
                                  const qb200_tensor* B, const int32_t* modesB, int32_t conjB, qb200_tensor* C,
                                  const int32_t* modesC, const double alpha[2], const double beta[2]) {
    using namespace qb;
    if (!ctx || !A || !B || !C) QB_FAIL(ctx, QB200_E_INVALID, "contract: null argument");
    const bool c64 = (A->dtype == QB200_C64);
    if ((A->dtype != QB200_C128 && !c64) || B->dtype != A->dtype || C->dtype != A->dtype)
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "contract: operands must all be ComplexF64 or all ComplexF32");
    ContractSpec spec;
    std::string err;
    int32_t r = make_contract_spec(A->rank, A->ext, modesA, B->rank, B->ext, modesB, C->rank, C->ext, modesC, &spec,
                                   &err);
    if (r != QB200_OK) QB_FAIL(ctx, r, "%s", err.c_str());
    Workspace ws(ctx);
    int64_t nent = contract_table_entries(spec);
    int64_t* tables = ws.get<int64_t>((size_t)nent);
    if (!tables) QB_FAIL(ctx, QB200_E_CUDA, "contract: workspace allocation failed");
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    QB_TRY(materialize_contract(ctx, spec, tables, &g));
    g.A = (const c128*)(spec.swapped ? B->data : A->data);
    g.B = (const c128*)(spec.swapped ? A->data : B->data);
    g.C = (c128*)C->data;
    g.conjA = spec.swapped ? (conjB != 0) : (conjA != 0);
    g.conjB = spec.swapped ? (conjA != 0) : (conjB != 0);
    g.alpha = alpha ? make_double2(alpha[0], alpha[1]) : make_double2(1.0, 0.0);
    g.beta = beta ? make_double2(beta[0], beta[1]) : make_double2(0.0, 0.0);
    g.beta_zero = (g.beta.x == 0.0 && g.beta.y == 0.0);
    return c64 ? launch_gemm_c64(ctx, g) : launch_gemm(ctx, g);
}
