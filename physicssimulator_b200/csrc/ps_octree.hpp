// ps_octree.hpp — host side of the SpatialTree candidate order (BroadPhaseCollisionDetectionKind.SpatialTree).
//
// The reference's second broad phase is a persistent 2^n-tree (Utilities/SpatialTree.fs) whose candidate list is
// history dependent: leaves split when they fill up and never merge again, every step removes the dynamic ids from
// all leaves and re-inserts them in ascending id order, and the candidates are the leaf-sharing pairs in depth-first
// leaf order (Collisions/BroadPhase.fs:62-70, 100-127).  The device walks whatever ORDERED candidate list it is given
// (the wavefront levels are built from the list, ps_step_kernel.cuh), so this mode is: extents from the device ->
// this tree on the host -> ordered list back to the device, once per step.  Node storage is a flat pool; leaf id sets
// are sorted vectors (the reference's Set<Id> iterates ascending).
//
// Host-only C++ (no CUDA), also compiled by tests/host_harness against the reference's known-answer facts
// (PhysicsSimulator.Tests/UtilitiesTests.fs:159-276).  Citations: /root/reference/PhysicsSimulatorCore/.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <set>
#include <string>
#include <unordered_set>
#include <vector>

namespace pst {

// ObjectExtent per dimension (SpatialTree.fs:21-28): Size and Position of the MIN corner
template <int D>
struct Ext {
    double size[D];
    double pos[D];
};

// areIntersecting (SpatialTree.fs:39-41) on withPositionCentered (:26-28): strict overlap on every axis
template <int D>
inline bool intersecting(const Ext<D>& a, const Ext<D>& b)
{
    for (int k = 0; k < D; k++) {
        const double ca = a.pos[k] + a.size[k] / 2.0;
        const double cb = b.pos[k] + b.size[k] / 2.0;
        if (!((a.size[k] + b.size[k]) / 2.0 > std::fabs(cb - ca))) return false;
    }
    return true;
}

// getSubspaceByIndex / getNodePositionByIndex (SpatialTree.fs:69-86): halves of the parent; the LAST dimension takes
// bit 0 of the child index (pinned by UtilitiesTests.fs:253-276)
template <int D>
inline Ext<D> subspace(const Ext<D>& parent, int index)
{
    Ext<D> s;
    int rest = index;
    for (int k = D - 1; k >= 0; k--) {
        const double off = (rest % 2 == 0) ? 0.0 : parent.size[k] / 2.0;
        rest /= 2;
        s.size[k] = parent.size[k] / 2.0;
        s.pos[k] = parent.pos[k] + off;
    }
    return s;
}

template <int D>
class Tree {
public:
    struct Node {
        int kids = -1;         // index of the first of 2^D consecutive children; -1 = leaf
        std::vector<int> ids;  // leaf: object ids, ascending
    };
    static constexpr int kKids = 1 << D;

    Tree() { nodes_.emplace_back(); }
    // SpatialTree.init (:93-110); false where the reference throws (non-positive arguments)
    bool init(int max_leaf, int max_depth, const double* lo, const double* hi)
    {
        if (max_leaf < 1 || max_depth < 1) return false;
        max_leaf_ = max_leaf;
        max_depth_ = max_depth;
        for (int k = 0; k < D; k++) {  // SpaceExtent.init (:33-36)
            space_.size[k] = hi[k] - lo[k];
            space_.pos[k] = lo[k];
        }
        nodes_.clear();
        nodes_.emplace_back();
        return true;
    }

    // insert (:166-208).  ext(id) gives the CURRENT extent of any object (a split re-files every object of the
    // leaf by its current extent, :179-182).  false = the object lies outside the tree's space: `insert` throws
    // ArgumentException there (:171-172), `tryInsert` (:210-214) swallows it — the caller decides.
    template <class ExtFn>
    bool insert(int id, ExtFn&& ext)
    {
        const Ext<D> oe = ext(id);
        if (!intersecting(space_, oe)) return false;
        add(0, 1, space_, id, oe, ext);
        return true;
    }

    // would insert() accept this extent?  (the test insert starts with, :171-172)
    bool in_space(const Ext<D>& oe) const { return intersecting(space_, oe); }

    // removeAll (:122-135): the ids for which gone(id) holds leave every leaf; the tree keeps its shape
    template <class Pred>
    void remove_if(Pred&& gone)
    {
        for (Node& n : nodes_)
            if (n.kids < 0) n.ids.erase(std::remove_if(n.ids.begin(), n.ids.end(), gone), n.ids.end());
    }

    // getObjectBuckets (:217-224): non-empty leaves, depth first, children in index order
    template <class F>
    void for_each_bucket(F&& f) const { walk(0, f); }

    // test hooks (the reference's facts start from hand-built trees, UtilitiesTests.fs:14-18, 223-250)
    void set_root_leaf(const std::vector<int>& ids)
    {
        nodes_.clear();
        nodes_.emplace_back();
        nodes_[0].ids = ids;
        std::sort(nodes_[0].ids.begin(), nodes_[0].ids.end());
    }
    void set_root_children(const std::vector<std::vector<int>>& kids)
    {
        nodes_.clear();
        nodes_.emplace_back();
        nodes_[0].kids = 1;
        for (int s = 0; s < kKids; s++) {
            nodes_.emplace_back();
            if (s < (int)kids.size()) nodes_.back().ids = kids[s];
            std::sort(nodes_.back().ids.begin(), nodes_.back().ids.end());
        }
    }
    std::string dump() const
    {
        std::string s;
        dump_node(0, s);
        return s;
    }
    size_t num_nodes() const { return nodes_.size(); }

private:
    template <class ExtFn>
    void add(int ni, int depth, const Ext<D>& ext_node, int id, const Ext<D>& oe, ExtFn&& ext)
    {
        if (nodes_[ni].kids < 0) {
            std::vector<int>& ids = nodes_[ni].ids;
            if ((int)ids.size() < max_leaf_ || depth >= max_depth_) {  // :186-188
                auto it = std::lower_bound(ids.begin(), ids.end(), id);
                if (it == ids.end() || *it != id) ids.insert(it, id);  // Set.add
                return;
            }
            // splitNode (:175-185): every object of the leaf plus the new one goes to each child it overlaps;
            // the children are leaves whatever their size (no recursive split)
            std::vector<int> all = ids;
            auto it = std::lower_bound(all.begin(), all.end(), id);
            if (it == all.end() || *it != id) all.insert(it, id);
            const int first = (int)nodes_.size();
            nodes_.resize(nodes_.size() + kKids);  // invalidates `ids`
            nodes_[ni].kids = first;
            nodes_[ni].ids.clear();
            for (int s = 0; s < kKids; s++) {
                const Ext<D> sub = subspace(ext_node, s);
                for (int o : all)
                    if (intersecting(o == id ? oe : ext(o), sub)) nodes_[first + s].ids.push_back(o);
            }
            return;
        }
        for (int s = 0; s < kKids; s++) {  // :189-199
            const Ext<D> sub = subspace(ext_node, s);
            if (intersecting(sub, oe)) add(nodes_[ni].kids + s, depth + 1, sub, id, oe, ext);
        }
    }
    template <class F>
    void walk(int ni, F&& f) const
    {
        const Node& n = nodes_[ni];
        if (n.kids < 0) {
            if (!n.ids.empty()) f(n.ids);
            return;
        }
        for (int s = 0; s < kKids; s++) walk(n.kids + s, f);
    }
    void dump_node(int ni, std::string& s) const  // "L{a,b}" / "N[.. ..]" (the spelling of the octree tests)
    {
        const Node& n = nodes_[ni];
        if (n.kids < 0) {
            s += "L{";
            for (size_t k = 0; k < n.ids.size(); k++) s += (k ? "," : "") + std::to_string(n.ids[k]);
            s += "}";
            return;
        }
        s += "N[";
        for (int c = 0; c < kKids; c++) {
            if (c) s += " ";
            dump_node(n.kids + c, s);
        }
        s += "]";
    }

    std::vector<Node> nodes_;
    Ext<D> space_{};
    int max_leaf_ = 4, max_depth_ = 10;
};

// getCollisionCandidates (BroadPhase.fs:62-70) then canIntersect (CollisionResolver.fs:58-60, 67): buckets with more
// than one id, the first occurrence of each distinct bucket, its 2-subsets in lexicographic order, the first
// occurrence of each unordered pair, minus static-static pairs.  Output: i | j << 16 with i < j (ids of one world).
inline void candidates(const Tree<3>& t, const unsigned char* is_static, std::vector<uint32_t>& out)
{
    out.clear();
    std::set<std::vector<int>> seen_buckets;
    std::unordered_set<uint32_t> seen_pairs;
    t.for_each_bucket([&](const std::vector<int>& ids) {
        if (ids.size() <= 1) return;
        if (!seen_buckets.insert(ids).second) return;
        for (size_t a = 0; a < ids.size(); a++)
            for (size_t b = a + 1; b < ids.size(); b++) {
                const uint32_t key = (uint32_t)ids[a] | ((uint32_t)ids[b] << 16);
                if (!seen_pairs.insert(key).second) continue;
                if (is_static[ids[a]] && is_static[ids[b]]) continue;
                out.push_back(key);
            }
    });
}

}  // namespace pst
