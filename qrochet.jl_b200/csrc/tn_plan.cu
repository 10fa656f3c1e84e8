// Planner of the sliced tensor-network contraction (replaces, deterministically, what the reference gets from
// `transform!(tn, ContractSimplification())`, `einexpr(tn; optimizer = HyPar(...))` and `findslices(SizeScorer(), path;
// size)` -- examples/distributed.jl:29-46; KaHyPar / EinExprs are un-vendored and randomised, so the rules are ours):
//   simplify    : while some pair of connected tensors contracts to a result no larger than its larger operand, contract
//                 the first such pair in (position, position) order of the live list and restart the scan;
//   greedy(a/2) : repeatedly contract the connected pair minimising 2 size(out) - a (size(x) + size(y)); ties -> smaller
//                 output -> earlier pair; disconnected remainder: outer product of the two smallest;
//   reconfigure : sub-tree reconfiguration: walk the tree from the root; at each internal node take the frontier of up to
//                 8 sub-trees below it (repeatedly open the largest openable one), find by exhaustive dynamic programming
//                 over the 2^8 subsets the order with the fewest flops (ties: smaller largest intermediate) and adopt it
//                 when strictly better; rounds repeat until nothing improves (at most 32);
//   candidates  : {simplify on, off} x a in {2, 4}; each is sliced by findslices and the one with the fewest total flops
//                 (slices x per-slice + slice-invariant work) wins, the first on ties;
//   findslices  : while the largest intermediate exceeds max_elements, cut the index with the largest
//                 score = sum of the sizes of the nodes holding it; ties -> the index met first in post-order.
// Sizes and flop counts are exact integers (unsigned __int128, saturating at 2^100; Python ints in
// oracle/circuit.py::plan, the same algorithm), so both implementations take identical decisions.
#include "tn_plan.cuh"

#include <algorithm>
#include <numeric>

namespace qb {
namespace {

const u128 CAP = (u128)1 << 100;
const u128 CAP_TOTAL = (u128)1 << 126;
constexpr int PLAN_K = 8;
constexpr int PLAN_ROUNDS = 32;
const int PLAN_ALPHAS[2] = {2, 4};

// Objective of the local search and of the candidate choice.  0: complex multiply-adds (EinExprs `flops`).  1: a time
// model of this executor on B200 in multiply-add units: a node costs max(macs, 10 x elements moved) -- the FP64 DMMA
// GEMM sustains ~4.1e12 complex MAC/s, HBM ~4.1e11 ComplexF64 elements/s, so a node with fewer than 10 MACs per element
// moved is bandwidth-bound.  Exact integers in both implementations.
inline u128 node_cost(int objective, u128 macs, u128 elems) {
    if (objective == 0) return macs;
    const u128 bw = elems > (((u128)1 << 110) / 10) ? ((u128)1 << 110) : elems * 10;
    return macs > bw ? macs : bw;
}

inline u128 sat_mul(u128 a, u128 b) {
    if (a == 0 || b == 0) return 0;
    if (a > CAP / b) return CAP;
    u128 r = a * b;
    return r > CAP ? CAP : r;
}

struct Net {
    std::vector<std::vector<int>> leaf_modes;
    std::vector<int64_t> ext;
    std::vector<int> total;
    int nm = 0;
    std::vector<char> mark;  // scratch

    u128 size(const std::vector<int>& ms, const std::vector<char>* cut = nullptr) const {
        u128 s = 1;
        for (int x : ms)
            if (!cut || !(*cut)[x]) s = sat_mul(s, (u128)ext[x]);
        return s;
    }
    PlanNode leaf(int i) const {
        PlanNode n;
        n.modes = leaf_modes[i];
        n.cnt.assign(nm, 0);
        for (int x : n.modes) n.cnt[x]++;
        return n;
    }
    // modes of the node contracting a and b, without building it
    void out_modes(const PlanNode& A, const PlanNode& B, std::vector<int>* out) {
        out->clear();
        for (int x : A.modes) {
            mark[x] = 1;
            if ((int)A.cnt[x] + (int)B.cnt[x] < total[x]) out->push_back(x);
        }
        for (int x : B.modes)
            if (!mark[x] && (int)A.cnt[x] + (int)B.cnt[x] < total[x]) out->push_back(x);
        for (int x : A.modes) mark[x] = 0;
    }
    PlanNode join(const std::vector<PlanNode>& nodes, int a, int b) {
        PlanNode n;
        n.left = a;
        n.right = b;
        out_modes(nodes[a], nodes[b], &n.modes);
        n.cnt.resize(nm);
        for (int x = 0; x < nm; ++x) n.cnt[x] = (uint16_t)(nodes[a].cnt[x] + nodes[b].cnt[x]);
        return n;
    }
    // complex multiply-adds of node i: product of the extents of every index involved (EinExprs `flops`)
    u128 flops(const std::vector<PlanNode>& nodes, int i, const std::vector<char>* cut = nullptr) {
        const PlanNode &A = nodes[nodes[i].left], &B = nodes[nodes[i].right];
        u128 s = 1;
        for (int x : A.modes) {
            mark[x] = 1;
            if (!cut || !(*cut)[x]) s = sat_mul(s, (u128)ext[x]);
        }
        for (int x : B.modes)
            if (!mark[x] && (!cut || !(*cut)[x])) s = sat_mul(s, (u128)ext[x]);
        for (int x : A.modes) mark[x] = 0;
        return s;
    }
    bool connected(const PlanNode& A, const PlanNode& B) {
        for (int x : A.modes) mark[x] = 1;
        bool c = false;
        for (int x : B.modes)
            if (mark[x]) {
                c = true;
                break;
            }
        for (int x : A.modes) mark[x] = 0;
        return c;
    }
};

void contract_pair(Net& net, std::vector<PlanNode>& nodes, std::vector<int>& live, int xi, int yi) {
    nodes.push_back(net.join(nodes, live[xi], live[yi]));
    live.erase(live.begin() + yi);
    live.erase(live.begin() + xi);
    live.push_back((int)nodes.size() - 1);
}

void simplify(Net& net, std::vector<PlanNode>& nodes, std::vector<int>& live) {
    std::vector<int> o;
    bool found = true;
    while (found) {
        found = false;
        for (size_t xi = 0; xi < live.size() && !found; ++xi)
            for (size_t yi = xi + 1; yi < live.size(); ++yi) {
                const PlanNode &A = nodes[live[xi]], &B = nodes[live[yi]];
                if (!net.connected(A, B)) continue;
                net.out_modes(A, B, &o);
                if (net.size(o) <= std::max(net.size(A.modes), net.size(B.modes))) {
                    contract_pair(net, nodes, live, (int)xi, (int)yi);
                    found = true;
                    break;
                }
            }
    }
}

void greedy(Net& net, std::vector<PlanNode>& nodes, std::vector<int>& live, int alpha2) {
    typedef __int128 i128;
    std::vector<int> o;
    std::vector<u128> sz;
    while (live.size() > 1) {
        sz.resize(live.size());
        for (size_t i = 0; i < live.size(); ++i) sz[i] = net.size(nodes[live[i]].modes);
        bool have = false;
        i128 best_cost = 0;
        u128 best_so = 0;
        int bx = -1, by = -1;
        for (size_t xi = 0; xi < live.size(); ++xi)
            for (size_t yi = xi + 1; yi < live.size(); ++yi) {
                const PlanNode &A = nodes[live[xi]], &B = nodes[live[yi]];
                if (!net.connected(A, B)) continue;
                net.out_modes(A, B, &o);
                const u128 so = net.size(o);
                const i128 cost = (i128)2 * (i128)so - (i128)alpha2 * (i128)(sz[xi] + sz[yi]);
                if (!have || cost < best_cost || (cost == best_cost && so < best_so)) {
                    have = true;
                    best_cost = cost;
                    best_so = so;
                    bx = (int)xi;
                    by = (int)yi;
                }
            }
        if (!have) {  // disconnected components: outer product of the two smallest
            std::vector<int> idx(live.size());
            std::iota(idx.begin(), idx.end(), 0);
            std::stable_sort(idx.begin(), idx.end(), [&](int p, int q) { return sz[p] < sz[q]; });
            bx = std::min(idx[0], idx[1]);
            by = std::max(idx[0], idx[1]);
        }
        contract_pair(net, nodes, live, bx, by);
    }
}

// One sub-tree reconfiguration at internal node i; true when the sub-tree was replaced by a cheaper one.
bool reconfigure_at(Net& net, std::vector<PlanNode>& nodes, int i, int objective) {
    std::vector<int> fr = {nodes[i].left, nodes[i].right};
    while ((int)fr.size() < PLAN_K) {
        int pick = -1;
        u128 pick_size = 0;
        for (size_t pos = 0; pos < fr.size(); ++pos) {
            const PlanNode& n = nodes[fr[pos]];
            if (n.left < 0) continue;
            u128 s = net.size(n.modes);
            if (pick < 0 || s > pick_size) {
                pick = (int)pos;
                pick_size = s;
            }
        }
        if (pick < 0) break;
        int j = fr[pick];
        fr.erase(fr.begin() + pick);
        fr.push_back(nodes[j].left);
        fr.push_back(nodes[j].right);
    }
    const int K = (int)fr.size();
    if (K < 3) return false;
    // cost of the current order above the frontier
    u128 old_fl = 0, old_mx = 0;
    {
        std::vector<int> stack = {i};
        while (!stack.empty()) {
            int j = stack.back();
            stack.pop_back();
            if (std::find(fr.begin(), fr.end(), j) != fr.end()) continue;
            old_fl += node_cost(objective, net.flops(nodes, j),
                                net.size(nodes[nodes[j].left].modes) + net.size(nodes[nodes[j].right].modes) + net.size(nodes[j].modes));
            old_mx = std::max(old_mx, net.size(nodes[j].modes));
            stack.push_back(nodes[j].left);
            stack.push_back(nodes[j].right);
        }
    }
    // relevant modes: the outer indices of the frontier sub-trees (everything else is summed inside one of them)
    std::vector<int> rel;
    for (int j : fr)
        for (int x : nodes[j].modes)
            if (!net.mark[x]) {
                net.mark[x] = 1;
                rel.push_back(x);
            }
    for (int x : rel) net.mark[x] = 0;
    const int R = (int)rel.size(), W = (R + 63) / 64;
    const int full = (1 << K) - 1;
    std::vector<uint16_t> cnt((size_t)(full + 1) * R, 0);
    std::vector<uint64_t> bits((size_t)(full + 1) * W, 0);
    std::vector<u128> size(full + 1, 1);
    auto set_size = [&](const uint64_t* b) {
        u128 s = 1;
        for (int w = 0; w < W; ++w) {
            uint64_t v = b[w];
            while (v) {
                int t = __builtin_ctzll(v);
                v &= v - 1;
                s = sat_mul(s, (u128)net.ext[rel[w * 64 + t]]);
            }
        }
        return s;
    };
    for (int m = 1; m <= full; ++m) {
        const int low = __builtin_ctz(m), rest = m & (m - 1);
        uint16_t* c = &cnt[(size_t)m * R];
        const PlanNode& n = nodes[fr[low]];
        for (int r = 0; r < R; ++r) c[r] = (uint16_t)(n.cnt[rel[r]] + (rest ? cnt[(size_t)rest * R + r] : 0));
        uint64_t* b = &bits[(size_t)m * W];
        for (int r = 0; r < R; ++r)
            if (c[r] > 0 && (int)c[r] < net.total[rel[r]]) b[r >> 6] |= (uint64_t)1 << (r & 63);
        size[m] = set_size(b);
    }
    struct Best {
        u128 fl = 0, mx = 0;
        int sub = 0, other = 0;
    };
    std::vector<Best> best(full + 1);
    std::vector<uint64_t> un(W);
    for (int m = 1; m <= full; ++m) {
        if ((m & (m - 1)) == 0) continue;
        const int low = m & -m;
        bool have = false;
        Best ch;
        for (int sub = (m - 1) & m; sub; sub = (sub - 1) & m) {
            if (!(sub & low)) continue;
            const int o = m ^ sub;
            for (int w = 0; w < W; ++w) un[w] = bits[(size_t)sub * W + w] | bits[(size_t)o * W + w];
            const u128 fl = best[sub].fl + best[o].fl + node_cost(objective, set_size(un.data()), size[sub] + size[o] + size[m]);
            const u128 mx = std::max(std::max(best[sub].mx, best[o].mx), size[m]);
            if (!have || fl < ch.fl || (fl == ch.fl && mx < ch.mx)) {
                have = true;
                ch.fl = fl;
                ch.mx = mx;
                ch.sub = sub;
                ch.other = o;
            }
        }
        best[m] = ch;
    }
    if (!(best[full].fl < old_fl || (best[full].fl == old_fl && best[full].mx < old_mx))) return false;
    struct Builder {
        Net& net;
        std::vector<PlanNode>& nodes;
        const std::vector<int>& fr;
        const std::vector<Best>& best;
        int full, root;
        int build(int m) {
            if ((m & (m - 1)) == 0) return fr[__builtin_ctz(m)];
            const int ia = build(best[m].sub);
            const int ib = build(best[m].other);
            PlanNode nd = net.join(nodes, ia, ib);
            if (m == full) {
                nodes[root] = nd;
                return root;
            }
            nodes.push_back(nd);
            return (int)nodes.size() - 1;
        }
    } builder{net, nodes, fr, best, full, i};
    builder.build(full);
    return true;
}

void reconfigure(Net& net, std::vector<PlanNode>& nodes, int root, int objective) {
    for (int round = 0; round < PLAN_ROUNDS; ++round) {
        bool improved = false;
        std::vector<int> stack = {root};
        while (!stack.empty()) {
            int i = stack.back();
            stack.pop_back();
            if (nodes[i].left < 0) continue;
            if (reconfigure_at(net, nodes, i, objective)) improved = true;
            stack.push_back(nodes[i].right);  // left sub-tree first
            stack.push_back(nodes[i].left);
        }
        if (!improved) break;
    }
}

// post-order renumbering of the reachable tree: leaves keep their ids, step s creates node nleaves + s
void compact(Net& net, const std::vector<PlanNode>& nodes, int root, std::vector<PlanNode>* out,
             std::vector<std::pair<int, int>>* path) {
    const int nleaves = (int)net.leaf_modes.size();
    std::vector<int> newid(nodes.size(), -1);
    path->clear();
    std::vector<std::pair<int, bool>> stack = {{root, false}};
    while (!stack.empty()) {
        auto [i, done] = stack.back();
        stack.pop_back();
        if (nodes[i].left < 0) {
            newid[i] = i;
        } else if (done) {
            path->push_back({newid[nodes[i].left], newid[nodes[i].right]});
            newid[i] = nleaves + (int)path->size() - 1;
        } else {
            stack.push_back({i, true});
            stack.push_back({nodes[i].right, false});
            stack.push_back({nodes[i].left, false});
        }
    }
    out->clear();
    for (int i = 0; i < nleaves; ++i) out->push_back(net.leaf(i));
    for (auto& p : *path) out->push_back(net.join(*out, p.first, p.second));
}

void findslices(Net& net, const std::vector<PlanNode>& nodes, int64_t max_elements, std::vector<int>* cut) {
    const int nleaves = (int)net.leaf_modes.size();
    cut->clear();
    if (max_elements <= 0) return;
    std::vector<char> is_cut(net.nm, 0);
    for (;;) {
        u128 mx = 0;
        for (size_t i = nleaves; i < nodes.size(); ++i) mx = std::max(mx, net.size(nodes[i].modes, &is_cut));
        if (mx <= (u128)max_elements) break;
        std::vector<u128> score(net.nm, 0);
        std::vector<char> seen_flag(net.nm, 0);
        std::vector<int> seen;
        for (size_t i = 0; i < nodes.size(); ++i) {  // compacted trees are stored in post-order
            const u128 s = net.size(nodes[i].modes, &is_cut);
            for (int x : nodes[i].modes) {
                if (is_cut[x] || net.ext[x] <= 1) continue;
                if (!seen_flag[x]) {
                    seen_flag[x] = 1;
                    seen.push_back(x);
                }
                score[x] += s;
            }
        }
        if (seen.empty()) break;
        int pick = seen[0];
        for (int x : seen)
            if (score[x] > score[pick]) pick = x;
        is_cut[pick] = 1;
        cut->push_back(pick);
    }
}

void sliced_cost(Net& net, const std::vector<PlanNode>& nodes, const std::vector<int>& cut, u128* per_slice, u128* once,
                 int64_t* nsl, int objective, u128* obj_per_slice, u128* obj_once) {
    const int nleaves = (int)net.leaf_modes.size();
    std::vector<char> is_cut(net.nm, 0);
    for (int x : cut) is_cut[x] = 1;
    std::vector<char> inv(nodes.size(), 1);
    for (int i = 0; i < nleaves; ++i)
        for (int x : net.leaf_modes[i])
            if (is_cut[x]) inv[i] = 0;
    *per_slice = *once = *obj_per_slice = *obj_once = 0;
    for (size_t i = nleaves; i < nodes.size(); ++i) {
        inv[i] = inv[nodes[i].left] && inv[nodes[i].right];
        const u128 f = net.flops(nodes, (int)i, &is_cut);
        const u128 c = node_cost(objective, f, net.size(nodes[nodes[i].left].modes, &is_cut) +
                                                  net.size(nodes[nodes[i].right].modes, &is_cut) + net.size(nodes[i].modes, &is_cut));
        if (inv[i]) {
            *once += f;
            *obj_once += c;
        } else {
            *per_slice += f;
            *obj_per_slice += c;
        }
    }
    int64_t n = 1;
    for (int x : cut) n = (n > ((int64_t)1 << 62) / net.ext[x]) ? ((int64_t)1 << 62) : n * net.ext[x];
    *nsl = n;
}

}  // namespace

PlanResult plan_network(const std::vector<std::vector<int>>& leaf_modes, const std::vector<int64_t>& ext,
                        int64_t max_elements, int optimizer) {
    Net net;
    net.leaf_modes = leaf_modes;
    net.ext = ext;
    net.nm = (int)ext.size();
    net.total.assign(net.nm, 0);
    net.mark.assign(net.nm, 0);
    for (auto& m : leaf_modes)
        for (int x : m) net.total[x]++;
    const int nleaves = (int)leaf_modes.size();
    PlanResult best;
    bool have = false;
    u128 best_total = 0;
    for (int simp = optimizer ? 1 : 0; simp >= 0; --simp)
        for (int ai = 0; ai < (optimizer ? 2 : 1); ++ai) {
            std::vector<PlanNode> nodes;
            for (int i = 0; i < nleaves; ++i) nodes.push_back(net.leaf(i));
            std::vector<int> live(nleaves);
            std::iota(live.begin(), live.end(), 0);
            if (simp) simplify(net, nodes, live);
            greedy(net, nodes, live, optimizer ? PLAN_ALPHAS[ai] : 2);
            const int root = live[0];
            const int objective = optimizer >= 2 ? 1 : 0;
            if (optimizer) reconfigure(net, nodes, root, objective);
            PlanResult r;
            compact(net, nodes, root, &r.nodes, &r.path);
            findslices(net, r.nodes, max_elements, &r.cut);
            u128 obj_ps = 0, obj_once = 0;
            sliced_cost(net, r.nodes, r.cut, &r.macs_per_slice, &r.macs_invariant, &r.nslices, objective, &obj_ps, &obj_once);
            u128 total;
            if (obj_ps != 0 && (u128)r.nslices > (CAP_TOTAL - std::min(CAP_TOTAL, obj_once)) / obj_ps)
                total = CAP_TOTAL;
            else
                total = std::min(CAP_TOTAL, obj_ps * (u128)r.nslices + obj_once);
            if (!have || total < best_total) {
                have = true;
                best_total = total;
                best = std::move(r);
            }
        }
    return best;
}

}  // namespace qb
