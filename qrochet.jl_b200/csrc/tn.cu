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
    // ---- executor state: built at the first contract call, rebuilt when the caller passes other leaf buffers ----
    struct LeafSlice {  // one leaf that holds cut indices: how its slice is gathered (device copy in leaf_desc)
        const c128* src;
        int64_t dst_off, nout;
        int32_t rank_out, ncut;
        int64_t ext_out[QB200_MAX_RANK], stride_out[QB200_MAX_RANK];
        int32_t cut_id[QB200_MAX_RANK];  // position in `sliced` of each cut index the leaf holds
        int64_t cut_stride[QB200_MAX_RANK];
    };
    bool exec_ready = false;
    bool exec_graph_mode = false;           // whether the executor was built with the slice graph enabled
    std::vector<const void*> bound_leaves;  // leaf data pointers the tables / graph were built for
    std::vector<c128*> node_ptr;            // every node: where its (sliced) data lives during the replay
    c128* arena = nullptr;                  // intermediates of the per-slice replay (offline first-fit over lifetimes)
    c128* inv_arena = nullptr;              // results of the slice-invariant nodes (recomputed once per call)
    c128* leaf_arena = nullptr;             // slices of the leaves that hold a cut index
    LeafSlice* leaf_desc = nullptr;         // device
    int nleaf_desc = 0;
    int64_t max_leaf_out = 0;
    int64_t* cut_meta = nullptr;            // device: [2 * ncut] = divisor, extent of each cut index
    int64_t* cursor = nullptr;              // device: [0] = slice to contract next, [1] = stride
    c128* acc = nullptr;                    // device accumulator of the rank-0 results
    cudaGraph_t graph = nullptr;            // one slice: gather the leaf slices, replay the tree, advance the cursor
    cudaGraphExec_t graph_exec = nullptr;
    int64_t launches_per_slice = 0;
    size_t arena_bytes = 0, inv_bytes = 0;
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
    std::set<int32_t> cut;
    for (int x : pr.cut) {
        cut.insert(label[x]);
        P->sliced.push_back(label[x]);
        P->sliced_ext.push_back(dext[x]);
    }
    // Storage order of an intermediate (free to choose: only the path and the cuts are part of the plan's contract):
    // the kept modes of the LARGER child first, in that child's order, then those of the smaller one.  The larger child
    // becomes the M side of the GEMM, whose rows enumerate its free modes in its own order, so both the reads of that
    // operand and the stores of the result walk memory contiguously (a gate-sized left child would otherwise put ITS
    // modes fastest and every 16-byte store of the big result would land in a sector of its own: measured 4x the
    // algorithmic L2 write traffic).
    auto sliced_size = [&](const TNNode& n) {
        double s = 1.0;
        for (size_t i = 0; i < n.modes.size(); ++i)
            if (!cut.count(n.modes[i])) s *= (double)n.ext[i];
        return s;
    };
    for (size_t id = ntensors; id < pr.nodes.size(); ++id) {
        TNNode n;
        n.left = pr.nodes[id].left;
        n.right = pr.nodes[id].right;
        std::set<int32_t> kept;
        for (int x : pr.nodes[id].modes) kept.insert(label[x]);
        const TNNode &L = P->nodes[n.left], &R = P->nodes[n.right];
        const bool right_first = sliced_size(R) > sliced_size(L);
        for (const TNNode* c : {right_first ? &R : &L, right_first ? &L : &R})
            for (size_t i = 0; i < c->modes.size(); ++i)
                if (kept.count(c->modes[i])) {
                    n.modes.push_back(c->modes[i]);
                    n.ext.push_back(c->ext[i]);
                    kept.erase(c->modes[i]);  // a kept shared (batch) mode is stored once, at its first position
                }
        P->nodes.push_back(n);
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
    *out = P;
    return QB200_OK;
}

static void release_executor(qb200_ctx* ctx, qb200_tnplan* P) {
    if (P->graph_exec) cudaGraphExecDestroy(P->graph_exec);
    if (P->graph) cudaGraphDestroy(P->graph);
    P->graph_exec = nullptr;
    P->graph = nullptr;
    if (ctx) {
        for (void* p : {(void*)P->arena, (void*)P->inv_arena, (void*)P->leaf_arena, (void*)P->leaf_desc, (void*)P->cut_meta,
                        (void*)P->cursor, (void*)P->acc})
            if (p) cudaFreeAsync(p, ctx->stream);
        for (auto& st : P->steps) {
            if (st.tables) cudaFreeAsync(st.tables, ctx->stream);
            st.tables = nullptr;
            st.ready = false;
        }
    }
    P->arena = P->inv_arena = P->leaf_arena = P->acc = nullptr;
    P->leaf_desc = nullptr;
    P->cut_meta = P->cursor = nullptr;
    P->exec_ready = false;
}

int32_t qb200_tn_plan_free(qb200_ctx* ctx, qb200_tnplan* P) {
    if (!P) return QB200_OK;
    release_executor(ctx, P);
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

// ---- executor ----------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// every leaf that holds a cut index, restricted to slice cursor[0] (first cut index fastest): ONE launch per slice
__global__ void tn_slice_leaves_kernel(const qb200_tnplan::LeafSlice* __restrict__ desc, const int64_t* __restrict__ cursor,
                                       const int64_t* __restrict__ cut_meta, c128* __restrict__ arena) {
    const qb200_tnplan::LeafSlice& L = desc[blockIdx.x];
    const int64_t s = cursor[0];
    int64_t base = 0;
    for (int j = 0; j < L.ncut; ++j) {
        const int64_t div = cut_meta[2 * L.cut_id[j]], ext = cut_meta[2 * L.cut_id[j] + 1];
        base += ((s / div) % ext) * L.cut_stride[j];
    }
    for (int64_t idx = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; idx < L.nout; idx += (int64_t)gridDim.y * blockDim.x) {
        int64_t rem = idx, off = base;
        for (int j = 0; j < L.rank_out; ++j) {
            off += (rem % L.ext_out[j]) * L.stride_out[j];
            rem /= L.ext_out[j];
        }
        arena[L.dst_off + idx] = L.src[off];
    }
}

__global__ void tn_advance_cursor_kernel(int64_t* cursor) { cursor[0] += cursor[1]; }

inline size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// offline first-fit allocator over the lifetimes of the replayed intermediates (one arena, offsets in bytes)
struct ArenaPlan {
    struct Block {
        size_t off, size;
    };
    std::vector<Block> free_list;
    size_t top = 0, peak = 0;  // the arena is sized by the PEAK: `top` shrinks again when trailing blocks are released
    size_t alloc(size_t bytes) {
        bytes = align_up(std::max<size_t>(bytes, 16));
        for (size_t i = 0; i < free_list.size(); ++i)
            if (free_list[i].size >= bytes) {
                size_t off = free_list[i].off;
                free_list[i].off += bytes;
                free_list[i].size -= bytes;
                if (free_list[i].size == 0) free_list.erase(free_list.begin() + i);
                return off;
            }
        size_t off = top;
        top += bytes;
        peak = std::max(peak, top);
        return off;
    }
    void release(size_t off, size_t bytes) {
        bytes = align_up(std::max<size_t>(bytes, 16));
        size_t i = 0;
        while (i < free_list.size() && free_list[i].off < off) ++i;
        free_list.insert(free_list.begin() + i, {off, bytes});
        if (i + 1 < free_list.size() && free_list[i].off + free_list[i].size == free_list[i + 1].off) {
            free_list[i].size += free_list[i + 1].size;
            free_list.erase(free_list.begin() + i + 1);
        }
        if (i > 0 && free_list[i - 1].off + free_list[i - 1].size == free_list[i].off) {
            free_list[i - 1].size += free_list[i].size;
            free_list.erase(free_list.begin() + i);
        }
        if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
            top = free_list.back().off;
            free_list.pop_back();
        }
    }
};

// a node is replayed per slice when it depends on a cut index; without cuts the whole tree is "the slice"
inline bool per_slice(const qb200_tnplan* P, int id) { return P->sliced.empty() || !P->nodes[id].invariant; }

int32_t launch_node(qb200_ctx* ctx, qb200_tnplan* P, int id) {
    const int nl = P->nleaves;
    const TNNode& n = P->nodes[id];
    TNStep& st = P->steps[id - nl];
    const bool last = (id == (int)P->nodes.size() - 1);
    GemmArgs g = st.args;
    g.A = st.spec.swapped ? P->node_ptr[n.right] : P->node_ptr[n.left];
    g.B = st.spec.swapped ? P->node_ptr[n.left] : P->node_ptr[n.right];
    g.C = last ? P->acc : P->node_ptr[id];  // the rank-0 result of every slice is accumulated on the device (beta = 1)
    g.conjA = g.conjB = 0;
    g.alpha = make_double2(1.0, 0.0);
    g.beta = last ? make_double2(1.0, 0.0) : make_double2(0.0, 0.0);
    g.beta_zero = last ? 0 : 1;
    PhaseTimer pt(ctx, QB_PH_TN_GEMM, 8.0 * g.M * (double)g.N * g.K * g.batch);
    return launch_gemm(ctx, g);
}

// one slice: gather the leaf slices at cursor[0], replay the per-slice part of the tree, advance the cursor
int32_t emit_slice(qb200_ctx* ctx, qb200_tnplan* P) {
    if (P->nleaf_desc > 0) {
        const unsigned chunks = (unsigned)std::min<int64_t>(std::max<int64_t>((P->max_leaf_out + 255) / 256, 1), 1024);
        tn_slice_leaves_kernel<<<dim3((unsigned)P->nleaf_desc, chunks), 256, 0, ctx->stream>>>(P->leaf_desc, P->cursor,
                                                                                              P->cut_meta, P->leaf_arena);
        QB_LAUNCH_CHECK(ctx);
    }
    static const bool debug_tn = getenv("QB200_DEBUG_TN") != nullptr;  // per-node timing (plain launches, one sync per node)
    for (int id = P->nleaves; id < (int)P->nodes.size(); ++id) {
        if (!per_slice(P, id)) continue;
        if (debug_tn) {
            double ms = 0.0;
            QB_TRY(qb200_timer_begin(ctx));
            QB_TRY(launch_node(ctx, P, id));
            QB_TRY(qb200_timer_end(ctx, &ms));
            const GemmArgs& g = P->steps[id - P->nleaves].args;
            const double macs = (double)g.M * g.N * g.K * g.batch;
            const double bytes = 16.0 * ((double)g.M * g.K + (double)g.K * g.N + (double)g.M * g.N) * g.batch;
            fprintf(stderr, "[qb200 tn] node %4d M %8d N %8d K %6d batch %5d tab(am ak bk bn cm cn) %d%d%d%d%d%d kfast %d%d  %8.1f us  %6.2f TF/s  %7.1f GB/s\n",
                    id, g.M, g.N, g.K, g.batch, g.am.tab != nullptr, g.ak.tab != nullptr, g.bk.tab != nullptr,
                    g.bn.tab != nullptr, g.cm.tab != nullptr, g.cn.tab != nullptr, g.a_kfast, g.b_kfast, ms * 1e3,
                    8.0 * macs / (ms * 1e9), bytes / (ms * 1e6));
        } else {
            QB_TRY(launch_node(ctx, P, id));
        }
    }
    tn_advance_cursor_kernel<<<1, 1, 0, ctx->stream>>>(P->cursor);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// QB200_TN_GRAPH=0 replays with plain launches (A/B and debugging); so do the phase profiler and QB200_SYNC_DEBUG
bool want_graph(qb200_ctx* ctx) {
    static const bool use_graph = [] {
        const char* e = getenv("QB200_TN_GRAPH");
        return !(e && e[0] == '0');
    }();
    return use_graph && !qb_sync_debug() && !ctx->prof_on && !getenv("QB200_DEBUG_TN");
}

// tables, arenas, leaf-slice descriptors and the slice graph for the given leaf buffers
int32_t build_executor(qb200_ctx* ctx, qb200_tnplan* P, qb200_tensor* const* leaves) {
    release_executor(ctx, P);
    const int nl = P->nleaves, nn = (int)P->nodes.size();
    std::set<int32_t> cut(P->sliced.begin(), P->sliced.end());
    auto sliced_elems = [&](const TNNode& n) {
        int64_t e = 1;
        for (size_t i = 0; i < n.modes.size(); ++i)
            if (!cut.count(n.modes[i])) e *= n.ext[i];
        return e;
    };
    P->node_ptr.assign(nn, nullptr);
    P->bound_leaves.assign(nl, nullptr);
    // leaves: invariant ones are read in place, the others are gathered per slice into the leaf arena
    std::vector<qb200_tnplan::LeafSlice> desc;
    std::vector<int> desc_leaf;
    size_t leaf_bytes = 0;
    P->max_leaf_out = 0;
    for (int t = 0; t < nl; ++t) {
        const TNNode& n = P->nodes[t];
        P->bound_leaves[t] = leaves[t]->data;
        if (n.invariant) {
            P->node_ptr[t] = (c128*)leaves[t]->data;
            continue;
        }
        qb200_tnplan::LeafSlice d;
        memset(&d, 0, sizeof(d));
        d.src = (const c128*)leaves[t]->data;
        d.dst_off = (int64_t)(leaf_bytes / sizeof(c128));
        d.nout = 1;
        int64_t stride = 1;
        for (size_t i = 0; i < n.modes.size(); ++i) {
            auto it = std::find(P->sliced.begin(), P->sliced.end(), n.modes[i]);
            if (it != P->sliced.end()) {
                d.cut_id[d.ncut] = (int32_t)(it - P->sliced.begin());
                d.cut_stride[d.ncut++] = stride;
            } else {
                d.ext_out[d.rank_out] = n.ext[i];
                d.stride_out[d.rank_out++] = stride;
                d.nout *= n.ext[i];
            }
            stride *= n.ext[i];
        }
        leaf_bytes += align_up((size_t)d.nout * sizeof(c128));
        P->max_leaf_out = std::max(P->max_leaf_out, d.nout);
        desc.push_back(d);
        desc_leaf.push_back(t);
    }
    P->nleaf_desc = (int)desc.size();
    // intermediates: slice-invariant nodes keep their own buffers, per-slice nodes share one arena by lifetime
    ArenaPlan ap;
    std::vector<size_t> off(nn, 0);
    size_t inv_bytes = 0;
    for (int id = nl; id < nn; ++id) {
        const TNNode& n = P->nodes[id];
        const size_t bytes = (size_t)sliced_elems(n) * sizeof(c128);
        if (id == nn - 1) continue;  // the root lands in the accumulator
        if (!per_slice(P, id)) {
            off[id] = inv_bytes;
            inv_bytes += align_up(std::max<size_t>(bytes, 16));
            continue;
        }
        off[id] = ap.alloc(bytes);
        for (int ch : {n.left, n.right})
            if (ch >= nl && per_slice(P, ch)) ap.release(off[ch], (size_t)sliced_elems(P->nodes[ch]) * sizeof(c128));
    }
    P->arena_bytes = ap.peak;
    P->inv_bytes = inv_bytes;
    auto dev_alloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(bytes, 256), ctx->stream) != cudaSuccess) return nullptr;
        return p;
    };
    P->arena = (c128*)dev_alloc(ap.peak);
    P->inv_arena = (c128*)dev_alloc(inv_bytes);
    P->leaf_arena = (c128*)dev_alloc(leaf_bytes);
    P->leaf_desc = (qb200_tnplan::LeafSlice*)dev_alloc(desc.size() * sizeof(qb200_tnplan::LeafSlice));
    P->cut_meta = (int64_t*)dev_alloc(2 * sizeof(int64_t) * std::max<size_t>(P->sliced.size(), 1));
    P->cursor = (int64_t*)dev_alloc(2 * sizeof(int64_t));
    P->acc = (c128*)dev_alloc(sizeof(c128));
    if (!P->arena || !P->inv_arena || !P->leaf_arena || !P->leaf_desc || !P->cut_meta || !P->cursor || !P->acc) {
        release_executor(ctx, P);
        QB_FAIL(ctx, QB200_E_CUDA, "tn_contract_sliced: out of device memory (arena %zu + %zu bytes)", ap.peak, inv_bytes);
    }
    for (size_t i = 0; i < desc.size(); ++i) P->node_ptr[desc_leaf[i]] = P->leaf_arena + desc[i].dst_off;
    for (int id = nl; id < nn - 1; ++id)
        P->node_ptr[id] = per_slice(P, id) ? (c128*)((char*)P->arena + off[id]) : (c128*)((char*)P->inv_arena + off[id]);
    std::vector<int64_t> meta(2 * std::max<size_t>(P->sliced.size(), 1), 1);
    {
        int64_t div = 1;
        for (size_t i = 0; i < P->sliced.size(); ++i) {
            meta[2 * i] = div;
            meta[2 * i + 1] = P->sliced_ext[i];
            div *= P->sliced_ext[i];
        }
    }
    if (!desc.empty())
        QB_CUDA(ctx, cudaMemcpyAsync(P->leaf_desc, desc.data(), desc.size() * sizeof(qb200_tnplan::LeafSlice),
                                     cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, cudaMemcpyAsync(P->cut_meta, meta.data(), meta.size() * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // desc / meta are host temporaries
    // offset tables of every step (device-built, kept for the life of the executor)
    for (int id = nl; id < nn; ++id) {
        TNStep& st = P->steps[id - nl];
        int64_t nent = contract_table_entries(st.spec);
        if (nent > 0) {
            st.tables = (int64_t*)dev_alloc(sizeof(int64_t) * nent);
            if (!st.tables) {
                release_executor(ctx, P);
                QB_FAIL(ctx, QB200_E_CUDA, "tn_contract_sliced: out of device memory (offset tables)");
            }
        }
        memset(&st.args, 0, sizeof(st.args));
        QB_TRY(materialize_contract(ctx, st.spec, st.tables, &st.args));
        st.ready = true;
    }
    // the slice as a CUDA graph: (1 + #per-slice nodes + 1) launches become one graph launch per slice
    P->exec_graph_mode = want_graph(ctx);
    if (P->exec_graph_mode) {
        const int64_t l0 = ctx->launches;
        QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        QB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int32_t r = emit_slice(ctx, P);
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        P->launches_per_slice = ctx->launches - l0;
        ctx->launches = l0;  // nothing ran yet
        if (r != QB200_OK || e != cudaSuccess || !g) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            release_executor(ctx, P);
            if (r != QB200_OK) return r;
            QB_FAIL(ctx, QB200_E_CUDA, "tn_contract_sliced: graph capture failed: %s", cudaGetErrorString(e));
        }
        P->graph = g;
        e = cudaGraphInstantiate(&P->graph_exec, g, 0);
        if (e != cudaSuccess) {
            release_executor(ctx, P);
            QB_FAIL(ctx, QB200_E_CUDA, "tn_contract_sliced: cudaGraphInstantiate: %s", cudaGetErrorString(e));
        }
    }
    P->exec_ready = true;
    return QB200_OK;
}

}  // namespace

extern "C" {

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
    if (nn == nl) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "tn_contract_sliced: network with a single tensor");
    bool rebuild = !P->exec_ready || P->exec_graph_mode != want_graph(ctx);
    for (int t = 0; t < nl && !rebuild; ++t) rebuild = (P->bound_leaves[t] != leaves[t]->data);
    if (rebuild) QB_TRY(build_executor(ctx, P, leaves));
    QB_CUDA(ctx, cudaMemsetAsync(P->acc, 0, sizeof(c128), ctx->stream));
    // slice-invariant sub-trees: once PER CALL (the leaf contents belong to the caller and may have changed)
    if (!P->sliced.empty())
        for (int id = nl; id < nn; ++id)
            if (P->nodes[id].invariant) QB_TRY(launch_node(ctx, P, id));
    int64_t* cur = reinterpret_cast<int64_t*>(ctx->scratch_host);
    cur[0] = first_slice;
    cur[1] = stride;
    QB_CUDA(ctx, cudaMemcpyAsync(P->cursor, cur, 2 * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    for (int64_t s = first_slice; s < P->nslices; s += stride) {
        if (P->graph_exec) {
            QB_CUDA(ctx, cudaGraphLaunch(P->graph_exec, ctx->stream));
            ctx->launches += P->launches_per_slice;
        } else {
            QB_TRY(emit_slice(ctx, P));
        }
    }
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // also: scratch_host (cursor) is free again
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, P->acc, sizeof(c128), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    acc[0] += ctx->scratch_host[0];
    acc[1] += ctx->scratch_host[1];
    return QB200_OK;
}

}  // extern "C"
