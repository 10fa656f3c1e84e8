// Contraction-tree planner of the sliced executor (host only): ContractSimplification + multi-start greedy + sub-tree
// reconfiguration + findslices.  See tn_plan.cu for the rules; oracle/circuit.py::plan is the same algorithm in Python
// and the tests compare the two decision by decision.
#pragma once
#include <stdint.h>

#include <utility>
#include <vector>

namespace qb {

typedef unsigned __int128 u128;

struct PlanNode {
    int left = -1, right = -1;
    std::vector<int> modes;     // dense mode ids, in the output order of the contraction
    std::vector<uint16_t> cnt;  // per dense mode: leaves below this node that hold it
};

struct PlanResult {
    std::vector<PlanNode> nodes;            // leaves, then one node per step (post-order)
    std::vector<std::pair<int, int>> path;  // step s contracts nodes path[s] into node nleaves + s
    std::vector<int> cut;                   // sliced modes (dense ids) in the order they were chosen
    u128 macs_per_slice = 0;                // complex multiply-adds of the nodes that depend on a cut mode, per slice
    u128 macs_invariant = 0;                // ... of the slice-invariant nodes (executed once)
    int64_t nslices = 1;
};

// leaf_modes: dense ids 0 .. nmodes-1; ext[mode]; optimizer 0 = one greedy tree (the round-1 rule), 1 = full planner
// minimising flops, 2 = full planner minimising the executor's time model (max(macs, 10 x elements moved) per node)
PlanResult plan_network(const std::vector<std::vector<int>>& leaf_modes, const std::vector<int64_t>& ext,
                        int64_t max_elements, int optimizer);

}  // namespace qb
