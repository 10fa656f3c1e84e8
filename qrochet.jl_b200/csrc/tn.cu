// K12: sliced contraction of a closed tensor network (examples/distributed.jl:29-101).
//
// Planner: tn_plan.cu (host C++, deterministic; ContractSimplification + multi-start greedy + sub-tree reconfiguration +
// findslices(SizeScorer) -- replaces EinExprs' Greedy/HyPar + findslices, whose own tie-breaks are randomised).
// Executor: for slice s (first cut index fastest) every leaf holding a cut index is restricted with
// select kernels, the fixed tree is replayed with the permutation-fused GEMM (offset tables are built
// once per plan and kept on the device), sub-trees that hold no cut index are contracted once per call and
// reused by every slice of that call, and the final rank-0 node is accumulated on the device (beta = 1).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <set>

#include "contract.cuh"
#include "tn_plan.cuh"

using namespace qb;

struct TNNode {
    int left = -1, right = -1;  // children ids (-1 for leaves)
    std::vector<int32_t> modes;
    std::vector<int64_t> ext;
    bool invariant = false;  // holds no cut index in its sub-tree
    int64_t size() const {
        int64_t s = 1;
        for (auto e : ext) s *= e;
        return s;
    }
};

struct TNStep {
    ContractSpec spec;
    GemmArgs args;          // table pointers filled at first execution
    int64_t* tables = nullptr;
    bool ready = false;
};

struct qb200_tnplan {
    int nleaves = 0;
    std::vector<TNNode> nodes;  // leaves then one node per step
    std::vector<int32_t> sliced;
    std::vector<int64_t> sliced_ext;
    int64_t nslices = 1;
    double flops_per_slice = 0.0;    // 8 x complex MACs of the nodes that depend on a cut index (executed per slice)
    double flops_invariant = 0.0;    // 8 x complex MACs of the slice-invariant nodes (executed once per call)
    int64_t max_inter = 0;
    std::vector<TNStep> steps;
    std::vector<c128*> cached;  // results of the slice-invariant nodes (device): valid inside ONE contract call only
};

namespace {

int64_t ext_of(const TNNode& n, int32_t mode) {
    for (size_t i = 0; i < n.modes.size(); ++i)
        if (n.modes[i] == mode) return n.ext[i];
    return -1;
}



}  // namespace

extern "C" {

int32_t qb200_tn_plan(qb200_ctx* ctx, int32_t ntensors, const int32_t* ranks, const int32_t* modes,
                      const int64_t* extents, int64_t max_elements, qb200_tnplan** out) {
    return qb200_tn_plan_opt(ctx, ntensors, ranks, modes, extents, max_elements, 1, out);
}

int32_t qb200_tn_plan_opt(qb200_ctx* ctx, int32_t ntensors, const int32_t* ranks, const int32_t* modes,
                          const int64_t* extents, int64_t max_elements, int32_t optimizer, qb200_tnplan** out) {
    // planning is pure host work: ctx may be NULL (then there is no error string, only the code)
    if (ntensors < 1 || !ranks || !modes || !extents || !out) QB_FAIL(ctx, QB200_E_INVALID, "tn_plan: bad argument");
    qb200_tnplan* P = new qb200_tnplan();
    P->nleaves = ntensors;
    std::map<int32_t, int> count;      // live tensors holding each index
    std::map<int32_t, int> count_all;  // over leaves only (open = 1)
    size_t off = 0;
    for (int t = 0; t < ntensors; ++t) {
        TNNode n;
        if (ranks[t] < 0 || ranks[t] > QB200_MAX_RANK) {
            delete P;
            QB_FAIL(ctx, QB200_E_INVALID, "tn_plan: bad leaf rank");
        }
        for (int i = 0; i < ranks[t]; ++i) {
            n.modes.push_back(modes[off + i]);
            n.ext.push_back(extents[off + i]);
            count[modes[off + i]]++;
        }
        off += ranks[t];
        P->nodes.push_back(n);
    }
    count_all = count;
    for (auto& kv : count_all)
        if (kv.second == 1) {
            delete P;
            QB_FAIL(ctx, QB200_E_UNSUPPORTED, "tn_plan: open index %d -- only closed networks (scalar result) are supported", kv.first);
        }
    for (auto& kv : count) {
        int64_t e = -1;
        for (auto& n : P->nodes) {
            int64_t x = ext_of(n, kv.first);
            if (x < 0) continue;
            if (e >= 0 && x != e) {
                delete P;
                QB_FAIL(ctx, QB200_E_INVALID, "tn_plan: index %d has inconsistent extents", kv.first);
            }
            e = x;
        }
    }

    // ---- path + slicing: tn_plan.cu on dense mode ids (numbered in order of first appearance) ----
    std::map<int32_t, int> dense;
    std::vector<int32_t> label;
    std::vector<int64_t> dext;
    std::vector<std::vector<int>> leaf_modes(ntensors);
    for (int t = 0; t < ntensors; ++t)
        for (size_t i = 0; i < P->nodes[t].modes.size(); ++i) {
            int32_t m = P->nodes[t].modes[i];
            auto it = dense.find(m);
            if (it == dense.end()) {
                it = dense.emplace(m, (int)label.size()).first;
                label.push_back(m);
                dext.push_back(P->nodes[t].ext[i]);
            }
            leaf_modes[t].push_back(it->second);
        }
    for (auto& lm : leaf_modes) {  // a label twice on one tensor (a trace) is not a pairwise-contraction network
        std::vector<int> srt = lm;
        std::sort(srt.begin(), srt.end());
        if (std::adjacent_find(srt.begin(), srt.end()) != srt.end()) {
            delete P;
            QB_FAIL(ctx, QB200_E_UNSUPPORTED, "tn_plan: repeated index on one tensor");
        }
    }
    PlanResult pr = plan_network(leaf_modes, dext, max_elements, optimizer);
    for (size_t id = ntensors; id < pr.nodes.size(); ++id) {
        TNNode n;
        n.left = pr.nodes[id].left;
        n.right = pr.nodes[id].right;
        for (int x : pr.nodes[id].modes) {
            n.modes.push_back(label[x]);
            n.ext.push_back(dext[x]);
        }
        P->nodes.push_back(n);
    }
    std::set<int32_t> cut;
    for (int x : pr.cut) {
        cut.insert(label[x]);
        P->sliced.push_back(label[x]);
        P->sliced_ext.push_back(dext[x]);
    }
    P->nslices = pr.nslices;
    auto node_size = [&](const TNNode& n) {
        int64_t s = 1;
        for (size_t i = 0; i < n.modes.size(); ++i)
            if (!cut.count(n.modes[i])) s *= n.ext[i];
        return s;
    };
    // ---- statistics and per-step contraction specs on the sliced shapes ----
    P->max_inter = 0;
    P->flops_per_slice = 0.0;
    // invariant flags
    for (size_t id = 0; id < P->nodes.size(); ++id) {
        TNNode& n = P->nodes[id];
        if (n.left < 0) {
            n.invariant = true;
            for (auto m : n.modes)
                if (cut.count(m)) n.invariant = false;
        } else {
            n.invariant = P->nodes[n.left].invariant && P->nodes[n.right].invariant;
        }
    }
    auto sliced_view = [&](const TNNode& n, std::vector<int32_t>* m, std::vector<int64_t>* e) {
        m->clear();
        e->clear();
        for (size_t i = 0; i < n.modes.size(); ++i)
            if (!cut.count(n.modes[i])) {
                m->push_back(n.modes[i]);
                e->push_back(n.ext[i]);
            }
    };
    P->steps.resize(P->nodes.size() - ntensors);
    for (size_t id = ntensors; id < P->nodes.size(); ++id) {
        const TNNode& n = P->nodes[id];
        std::vector<int32_t> ma, mb, mc;
        std::vector<int64_t> ea, eb, ec;
        sliced_view(P->nodes[n.left], &ma, &ea);
        sliced_view(P->nodes[n.right], &mb, &eb);
        sliced_view(n, &mc, &ec);
        std::string err;
        int32_t r = make_contract_spec((int)ma.size(), ea.data(), ma.data(), (int)mb.size(), eb.data(), mb.data(),
                                       (int)mc.size(), ec.data(), mc.data(), &P->steps[id - ntensors].spec, &err);
        if (r != QB200_OK) {
            delete P;
            QB_FAIL(ctx, r, "tn_plan: step %zu: %s", id - ntensors, err.c_str());
        }
        std::set<int32_t> involved(ma.begin(), ma.end());
        involved.insert(mb.begin(), mb.end());
        double macs = 1.0;
        for (auto m : involved) {
            int64_t e = ext_of(P->nodes[n.left], m);
            if (e < 0) e = ext_of(P->nodes[n.right], m);
            macs *= (double)e;
        }
        (n.invariant ? P->flops_invariant : P->flops_per_slice) += 8.0 * macs;
        P->max_inter = std::max(P->max_inter, node_size(n));
    }
    P->cached.assign(P->nodes.size(), nullptr);
    *out = P;
    return QB200_OK;
}

int32_t qb200_tn_plan_free(qb200_ctx* ctx, qb200_tnplan* P) {
    if (!P) return QB200_OK;
    for (auto& s : P->steps)
        if (s.tables && ctx) cudaFreeAsync(s.tables, ctx->stream);
    for (auto p : P->cached)
        if (p && ctx) cudaFreeAsync(p, ctx->stream);
    delete P;
    return QB200_OK;
}

int64_t qb200_tn_plan_nslices(const qb200_tnplan* P) { return P ? P->nslices : -1; }
int32_t qb200_tn_plan_sliced_modes(const qb200_tnplan* P, int32_t* modes_out) {
    if (!P) return -1;
    if (modes_out)
        for (size_t i = 0; i < P->sliced.size(); ++i) modes_out[i] = P->sliced[i];
    return (int32_t)P->sliced.size();
}
double qb200_tn_plan_flops_per_slice(const qb200_tnplan* P) { return P ? P->flops_per_slice : -1.0; }
double qb200_tn_plan_flops_invariant(const qb200_tnplan* P) { return P ? P->flops_invariant : -1.0; }
int64_t qb200_tn_plan_max_intermediate(const qb200_tnplan* P) { return P ? P->max_inter : -1; }
int32_t qb200_tn_plan_path(const qb200_tnplan* P, int32_t* pairs_out) {
    if (!P) return -1;
    int nsteps = (int)P->nodes.size() - P->nleaves;
    if (pairs_out)
        for (int s = 0; s < nsteps; ++s) {
            pairs_out[2 * s] = P->nodes[P->nleaves + s].left;
            pairs_out[2 * s + 1] = P->nodes[P->nleaves + s].right;
        }
    return nsteps;
}

int32_t qb200_tn_contract_sliced(qb200_ctx* ctx, qb200_tnplan* P, qb200_tensor* const* leaves, int64_t first_slice,
                                 int64_t stride, double acc[2]) {
    if (!ctx || !P || !leaves || !acc || stride < 1 || first_slice < 0) QB_FAIL(ctx, QB200_E_INVALID, "tn_contract_sliced: bad argument");
    const int nl = P->nleaves, nn = (int)P->nodes.size();
    for (int t = 0; t < nl; ++t) {
        const qb200_tensor* L = leaves[t];
        if (!L || L->dtype != QB200_C128 || L->rank != (int)P->nodes[t].modes.size())
            QB_FAIL(ctx, QB200_E_INVALID, "tn_contract_sliced: leaf %d does not match the plan", t);
        for (int i = 0; i < L->rank; ++i)
            if (L->ext[i] != P->nodes[t].ext[i]) QB_FAIL(ctx, QB200_E_INVALID, "tn_contract_sliced: leaf %d extent mismatch", t);
    }
    Workspace ws(ctx);
    c128* accd = ws.get<c128>(1);
    if (!accd) QB_FAIL(ctx, QB200_E_CUDA, "tn_contract_sliced: workspace allocation failed");
    QB_CUDA(ctx, cudaMemsetAsync(accd, 0, sizeof(c128), ctx->stream));
    if (nn == nl) {
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "tn_contract_sliced: network with a single tensor");
    }
    std::set<int32_t> cut(P->sliced.begin(), P->sliced.end());
    // slice-invariant sub-trees are contracted once PER CALL: `leaves` belong to the caller and may hold other data
    // the next time (a cache that outlived the call returned stale sub-trees)
    auto drop_cache = [&]() {
        for (auto& p : P->cached)
            if (p) {
                cudaFreeAsync(p, ctx->stream);
                p = nullptr;
            }
    };
    drop_cache();
    std::vector<c128*> buf(nn, nullptr);
    std::vector<bool> owned(nn, false);
    const c128 ONE = {1.0, 0.0}, ZERO = {0.0, 0.0};

    for (int64_t s = first_slice; s < P->nslices; s += stride) {
        // cut index values, first cut index fastest (Iterators.product order, examples/distributed.jl:47,69)
        std::map<int32_t, int64_t> value;
        {
            int64_t rem = s;
            for (size_t i = 0; i < P->sliced.size(); ++i) {
                value[P->sliced[i]] = rem % P->sliced_ext[i];
                rem /= P->sliced_ext[i];
            }
        }
        // leaves restricted to this slice
        for (int t = 0; t < nl; ++t) {
            const TNNode& n = P->nodes[t];
            if (n.invariant) {
                buf[t] = (c128*)leaves[t]->data;
                owned[t] = false;
                continue;
            }
            qb200_tensor cur = *leaves[t];
            cur.owned = false;
            std::vector<int32_t> m = n.modes;
            c128* held = nullptr;
            for (size_t i = 0; i < m.size();) {
                if (!cut.count(m[i])) {
                    ++i;
                    continue;
                }
                qb200_tensor nxt = cur;
                nxt.rank = cur.rank - 1;
                int64_t cnt = 1;
                for (int j = 0, k2 = 0; j < cur.rank; ++j)
                    if (j != (int)i) {
                        nxt.ext[k2++] = cur.ext[j];
                        cnt *= cur.ext[j];
                    }
                void* p = nullptr;
                QB_CUDA(ctx, cudaMallocAsync(&p, sizeof(c128) * std::max<int64_t>(cnt, 1), ctx->stream));
                nxt.data = p;
                int32_t r = qb200_select_mode(ctx, &cur, (int32_t)i, value[m[i]], &nxt);
                if (held) cudaFreeAsync(held, ctx->stream);
                held = (c128*)p;
                if (r != QB200_OK) {
                    cudaFreeAsync(held, ctx->stream);
                    return r;
                }
                cur = nxt;
                m.erase(m.begin() + i);
            }
            buf[t] = (c128*)cur.data;
            owned[t] = true;
        }
        // replay the tree
        for (int id = nl; id < nn; ++id) {
            TNNode& n = P->nodes[id];
            TNStep& st = P->steps[id - nl];
            const bool last = (id == nn - 1);
            if (n.invariant && P->cached[id]) {
                buf[id] = P->cached[id];
                owned[id] = false;
                continue;
            }
            if (!st.ready) {
                int64_t nent = contract_table_entries(st.spec);
                if (nent > 0) {
                    void* p = nullptr;
                    QB_CUDA(ctx, cudaMallocAsync(&p, sizeof(int64_t) * nent, ctx->stream));
                    st.tables = (int64_t*)p;
                }
                memset(&st.args, 0, sizeof(st.args));
                QB_TRY(materialize_contract(ctx, st.spec, st.tables, &st.args));
                st.ready = true;
            }
            int64_t osize = 1;
            for (size_t i = 0; i < n.modes.size(); ++i)
                if (!cut.count(n.modes[i])) osize *= n.ext[i];
            c128* outp;
            if (last) {
                outp = accd;  // rank-0 result accumulated on the device
            } else {
                void* p = nullptr;
                cudaError_t e = cudaMallocAsync(&p, sizeof(c128) * osize, ctx->stream);
                if (e != cudaSuccess) QB_FAIL(ctx, QB200_E_CUDA, "tn_contract_sliced: out of device memory (%lld elements)", (long long)osize);
                outp = (c128*)p;
            }
            GemmArgs g = st.args;
            g.A = st.spec.swapped ? buf[n.right] : buf[n.left];
            g.B = st.spec.swapped ? buf[n.left] : buf[n.right];
            g.C = outp;
            g.conjA = g.conjB = 0;
            g.alpha = ONE;
            g.beta = last ? ONE : ZERO;
            g.beta_zero = last ? 0 : 1;
            if (getenv("QB200_DEBUG_TN") && s == first_slice) {
                cudaEvent_t e0, e1;
                cudaEventCreate(&e0);
                cudaEventCreate(&e1);
                cudaEventRecord(e0, ctx->stream);
                QB_TRY(launch_gemm(ctx, g));
                cudaEventRecord(e1, ctx->stream);
                cudaEventSynchronize(e1);
                float msf = 0;
                cudaEventElapsedTime(&msf, e0, e1);
                double fl = 8.0 * g.M * (double)g.N * g.K * g.batch;
                fprintf(stderr, "[tn] step %d M %d N %d K %d batch %d akfast %d bkfast %d tabs %d%d%d%d%d%d  %.3f ms %.2f TF/s inv %d\n",
                        id - nl, g.M, g.N, g.K, g.batch, g.a_kfast, g.b_kfast, g.am.tab != nullptr, g.ak.tab != nullptr,
                        g.bk.tab != nullptr, g.bn.tab != nullptr, g.cm.tab != nullptr, g.cn.tab != nullptr, msf,
                        fl / msf / 1e9, (int)n.invariant);
                cudaEventDestroy(e0);
                cudaEventDestroy(e1);
            } else
                QB_TRY(launch_gemm(ctx, g));
            // children are consumed exactly once in a tree
            for (int ch : {n.left, n.right}) {
                if (owned[ch] && buf[ch]) cudaFreeAsync(buf[ch], ctx->stream);
                buf[ch] = nullptr;
                owned[ch] = false;
            }
            if (last) {
                buf[id] = nullptr;
            } else if (n.invariant) {
                P->cached[id] = outp;  // reused by every later slice of this call
                buf[id] = outp;
                owned[id] = false;
            } else {
                buf[id] = outp;
                owned[id] = true;
            }
        }
    }
    drop_cache();
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, accd, sizeof(c128), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    acc[0] += ctx->scratch_host[0];
    acc[1] += ctx->scratch_host[1];
    return QB200_OK;
}

}  // extern "C"
