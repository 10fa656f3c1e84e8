// K12 + planner: sliced contraction of a closed tensor network (examples/distributed.jl:29-101).
//
// Planner (host C++, deterministic; replaces EinExprs' Greedy/HyPar + findslices(SizeScorer), whose own
// tie-breaks are randomised):
//   path    : greedy -- among all pairs of tensors sharing an index pick the one minimising
//             size(out) - size(a) - size(b); ties -> smaller output -> lower ids.  An index is summed
//             as soon as no other live tensor holds it (hyper-index aware).
//   slicing : while the largest intermediate exceeds max_elements, slice the index with the largest
//             score = sum of the sizes of all path nodes holding it; ties -> the index met first in a
//             post-order walk of the path.
// Executor: for slice s (first cut index fastest) every leaf holding a cut index is restricted with
// select kernels, the fixed tree is replayed with the permutation-fused GEMM (offset tables are built
// once per plan and kept on the device), sub-trees that hold no cut index are contracted once and
// reused by every slice, and the final rank-0 node is accumulated on the device (beta = 1).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <set>

#include "contract.cuh"

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
    double flops_per_slice = 0.0;
    int64_t max_inter = 0;
    std::vector<TNStep> steps;
    std::vector<c128*> cached;  // results of invariant nodes (device), filled on first use
    bool cache_valid = false;
};

namespace {

int64_t ext_of(const TNNode& n, int32_t mode) {
    for (size_t i = 0; i < n.modes.size(); ++i)
        if (n.modes[i] == mode) return n.ext[i];
    return -1;
}

bool holds(const TNNode& n, int32_t mode) { return std::find(n.modes.begin(), n.modes.end(), mode) != n.modes.end(); }

// output modes of contracting a and b when `count` tells how many live tensors hold each index
void out_modes(const TNNode& a, const TNNode& b, const std::map<int32_t, int>& count, TNNode* out) {
    out->modes.clear();
    out->ext.clear();
    for (size_t i = 0; i < a.modes.size(); ++i) {
        int32_t m = a.modes[i];
        int users = count.at(m) - 1 - (holds(b, m) ? 1 : 0);
        if (users > 0) {
            out->modes.push_back(m);
            out->ext.push_back(a.ext[i]);
        }
    }
    for (size_t i = 0; i < b.modes.size(); ++i) {
        int32_t m = b.modes[i];
        if (holds(a, m)) continue;
        int users = count.at(m) - 1;
        if (users > 0) {
            out->modes.push_back(m);
            out->ext.push_back(b.ext[i]);
        }
    }
}

void post_order(const std::vector<TNNode>& nodes, int id, std::vector<int>* order) {
    if (nodes[id].left >= 0) {
        post_order(nodes, nodes[id].left, order);
        post_order(nodes, nodes[id].right, order);
    }
    order->push_back(id);
}

}  // namespace

extern "C" {

int32_t qb200_tn_plan(qb200_ctx* ctx, int32_t ntensors, const int32_t* ranks, const int32_t* modes,
                      const int64_t* extents, int64_t max_elements, qb200_tnplan** out) {
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

    // ---- greedy path ----
    std::vector<int> live;
    for (int t = 0; t < ntensors; ++t) live.push_back(t);
    while (live.size() > 1) {
        bool found = false;
        int64_t best_cost = 0, best_size = 0;
        int bi = -1, bj = -1;
        TNNode best_out;
        for (size_t x = 0; x < live.size(); ++x)
            for (size_t y = x + 1; y < live.size(); ++y) {
                const TNNode& a = P->nodes[live[x]];
                const TNNode& b = P->nodes[live[y]];
                bool connected = false;
                for (auto m : a.modes)
                    if (holds(b, m)) {
                        connected = true;
                        break;
                    }
                if (!connected) continue;
                TNNode o;
                out_modes(a, b, count, &o);
                int64_t so = o.size();
                int64_t cost = so - a.size() - b.size();
                if (!found || cost < best_cost || (cost == best_cost && so < best_size)) {
                    found = true;
                    best_cost = cost;
                    best_size = so;
                    bi = (int)x;
                    bj = (int)y;
                    best_out = o;
                }
            }
        if (!found) {  // disconnected components: outer product of the two smallest
            std::vector<size_t> idx(live.size());
            for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
            std::stable_sort(idx.begin(), idx.end(),
                             [&](size_t p, size_t q) { return P->nodes[live[p]].size() < P->nodes[live[q]].size(); });
            bi = (int)std::min(idx[0], idx[1]);
            bj = (int)std::max(idx[0], idx[1]);
            out_modes(P->nodes[live[bi]], P->nodes[live[bj]], count, &best_out);
        }
        int ia = live[bi], ib = live[bj];
        for (auto m : P->nodes[ia].modes) count[m]--;
        for (auto m : P->nodes[ib].modes) count[m]--;
        for (auto m : best_out.modes) count[m]++;
        best_out.left = ia;
        best_out.right = ib;
        P->nodes.push_back(best_out);
        live.erase(live.begin() + bj);
        live.erase(live.begin() + bi);
        live.push_back((int)P->nodes.size() - 1);
    }

    // ---- findslices ----
    std::vector<int> order;
    post_order(P->nodes, (int)P->nodes.size() - 1, &order);
    std::set<int32_t> cut;
    auto node_size = [&](const TNNode& n) {
        int64_t s = 1;
        for (size_t i = 0; i < n.modes.size(); ++i)
            if (!cut.count(n.modes[i])) s *= n.ext[i];
        return s;
    };
    if (max_elements > 0) {
        for (;;) {
            int64_t mx = 0;
            for (size_t id = ntensors; id < P->nodes.size(); ++id) mx = std::max(mx, node_size(P->nodes[id]));
            if (mx <= max_elements) break;
            std::map<int32_t, double> score;
            std::vector<int32_t> first_seen;
            for (int id : order) {
                const TNNode& n = P->nodes[id];
                double s = (double)node_size(n);
                for (size_t i = 0; i < n.modes.size(); ++i) {
                    int32_t m = n.modes[i];
                    if (cut.count(m) || n.ext[i] <= 1) continue;
                    if (!score.count(m)) first_seen.push_back(m);
                    score[m] += s;
                }
            }
            if (first_seen.empty()) break;
            int32_t pick = first_seen[0];
            for (int32_t m : first_seen)
                if (score[m] > score[pick]) pick = m;
            cut.insert(pick);
            P->sliced.push_back(pick);
            int64_t e = 1;
            for (auto& n : P->nodes) {
                int64_t x = ext_of(n, pick);
                if (x > 0) e = x;
            }
            P->sliced_ext.push_back(e);
            P->nslices *= e;
        }
    }
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
        P->flops_per_slice += 8.0 * macs;
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
                P->cached[id] = outp;  // reused by every later slice (and later calls with the same leaves)
                buf[id] = outp;
                owned[id] = false;
            } else {
                buf[id] = outp;
                owned[id] = true;
            }
        }
    }
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, accd, sizeof(c128), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    acc[0] += ctx->scratch_host[0];
    acc[1] += ctx->scratch_host[1];
    return QB200_OK;
}

}  // extern "C"
