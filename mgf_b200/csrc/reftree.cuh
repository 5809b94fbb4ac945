// reftree.cuh -- BVH<AABB, V> exactly as src/bvh.rs grows and shrinks it, on the HOST (product code; used where the ORDER of the
// reference's query callbacks is part of the result: Compound, compound.cuh, and World::step in reference order, reforder.cuh).
//   insert  bvh.rs:125-217  descend by the surface-area heuristic, hang the new leaf beside the chosen node under a fresh parent,
//                           walk up refitting (rounded unions, bounds.rs:113-130) and rotating (balance, bvh.rs:371-480)
//   remove  bvh.rs:220-260  the sibling takes the parent's place, then the same walk up
//   query   bvh.rs:283-310  explicit stack, left child pushed first: the RIGHT child is visited first
// Slots are recycled last-freed-first like pool.rs:81-113 (only identities: the shape of the tree does not depend on them).
#pragma once
#include <vector>

namespace mgfb {

struct CompNode {          // BVHNode<AABB, V> (bvh.rs:32-46)
    float4 c, r;           // bounds
    int left, right;       // BVHNodeType::Parent(l, r); leaf: left = -1, right = the value
    int parent, height;
};
HD void box_combine(V3 ac, V3 ar, V3 bc, V3 br, V3* oc, V3* orr) {   // bounds.rs:113-130
    V3 lo = mk3(fminf(ac.x - ar.x, bc.x - br.x), fminf(ac.y - ar.y, bc.y - br.y), fminf(ac.z - ar.z, bc.z - br.z));
    V3 hi = mk3(fmaxf(ac.x + ar.x, bc.x + br.x), fmaxf(ac.y + ar.y, bc.y + br.y), fmaxf(ac.z + ar.z, bc.z + br.z));
    *orr = (hi - lo) / 2.0f; *oc = (hi + lo) / 2.0f;
}

struct RefTree {
    std::vector<CompNode> N;
    std::vector<int> free_slots;   // last freed on top
    int root = 0, count = 0;       // count = occupied slots (Pool::len)

    static V3 c(const CompNode& n) { return mk3(n.c.x, n.c.y, n.c.z); }
    static V3 r(const CompNode& n) { return mk3(n.r.x, n.r.y, n.r.z); }
    static float area(V3 rr) { return rr.x * rr.y + rr.y * rr.z + rr.z * rr.x; }   // bounds.rs:132-134
    bool is_leaf(int i) const { return N[i].left < 0; }
    bool empty() const { return count == 0; }
    void clear() { N.clear(); free_slots.clear(); root = 0; count = 0; }

    int add(V3 bc, V3 br, int left, int right) {   // insert_node + Pool::push
        CompNode n; n.c = make_float4(bc.x, bc.y, bc.z, 0.0f); n.r = make_float4(br.x, br.y, br.z, 0.0f);
        n.left = left; n.right = right; n.parent = 0; n.height = -1;
        ++count;
        if (!free_slots.empty()) { int s = free_slots.back(); free_slots.pop_back(); N[s] = n; return s; }
        N.push_back(n);
        return (int)N.size() - 1;
    }
    void drop(int i) { free_slots.push_back(i); --count; }   // Pool::remove
    bool refit(int i) {   // bounds = combine(left, right); false when the reference's assert!(r >= 0) would fire
        V3 oc, orr; box_combine(c(N[N[i].left]), r(N[N[i].left]), c(N[N[i].right]), r(N[N[i].right]), &oc, &orr);
        N[i].c = make_float4(oc.x, oc.y, oc.z, 0.0f); N[i].r = make_float4(orr.x, orr.y, orr.z, 0.0f);
        return orr.x >= 0.0f && orr.y >= 0.0f && orr.z >= 0.0f;
    }
    // one rotation of bvh.rs:371-480: `up` (the taller child of a) takes a's place; of up's children the taller stays with up
    // beside a, the other replaces up under a.  `up_is_right`: up was a's right child (then a keeps its left child first).
    int rotate(int a, int up, bool up_is_right) {
        const int keep = up_is_right ? N[a].left : N[a].right;      // a's other child
        const int x = N[up].left, y = N[up].right;
        N[up].parent = N[a].parent; N[a].parent = up;
        if (root == a) root = up;
        else if (!is_leaf(N[up].parent)) { CompNode& p = N[N[up].parent]; if (p.left == a) p.left = up; else p.right = up; }
        const int stay = N[x].height > N[y].height ? x : y, move = stay == x ? y : x;
        N[up].left = a; N[up].right = stay;
        if (up_is_right) { N[a].left = keep; N[a].right = move; } else { N[a].left = move; N[a].right = keep; }
        N[move].parent = a;
        // combine(b, g) / combine(c, e): the kept child is the first argument in both mirror cases
        { V3 oc, orr; box_combine(c(N[keep]), r(N[keep]), c(N[move]), r(N[move]), &oc, &orr); N[a].c = make_float4(oc.x, oc.y, oc.z, 0.0f); N[a].r = make_float4(orr.x, orr.y, orr.z, 0.0f); }
        { V3 oc, orr; box_combine(c(N[a]), r(N[a]), c(N[stay]), r(N[stay]), &oc, &orr); N[up].c = make_float4(oc.x, oc.y, oc.z, 0.0f); N[up].r = make_float4(orr.x, orr.y, orr.z, 0.0f); }
        N[a].height = 1 + std::max(N[keep].height, N[move].height);
        N[up].height = 1 + std::max(N[a].height, N[stay].height);
        return up;
    }
    int balance(int a) {
        if (N[a].height < 2 || is_leaf(a)) return a;
        const int b = N[a].left, cc = N[a].right;
        if (N[cc].height > N[b].height + 1) return is_leaf(cc) ? cc : rotate(a, cc, true);
        if (N[b].height > N[cc].height + 1) return is_leaf(b) ? b : rotate(a, b, false);
        return a;
    }
    // BVH::insert: returns the leaf's index, or -1 when a union's half extent goes negative / NaN (bounds.rs:125-127 asserts)
    int insert(V3 bc, V3 br, int value) {
        const int leaf = add(bc, br, -1, value);
        if (count == 1) { root = leaf; return leaf; }
        int best = root;
        while (!is_leaf(best)) {
            const int c1 = N[best].left, c2 = N[best].right;
            const float a0 = area(r(N[best]));
            V3 oc, orr; box_combine(c(N[best]), r(N[best]), bc, br, &oc, &orr);
            const float combined = area(orr);
            const float no_descent = combined * 2.0f, inherit = (combined - a0) * 2.0f;
            auto child_cost = [&](int ch) {
                V3 qc, qr; box_combine(bc, br, c(N[ch]), r(N[ch]), &qc, &qr);
                return is_leaf(ch) ? area(qr) + inherit : area(qr) - area(r(N[ch])) + inherit;
            };
            const float k1 = child_cost(c1), k2 = child_cost(c2);
            if (no_descent < k1 && no_descent < k2) break;
            best = k1 < k2 ? c1 : c2;
        }
        const int old_parent = N[best].parent;
        V3 oc, orr; box_combine(bc, br, c(N[best]), r(N[best]), &oc, &orr);
        if (!(orr.x >= 0.0f && orr.y >= 0.0f && orr.z >= 0.0f)) return -1;
        const int np = add(oc, orr, best, leaf);
        N[np].parent = old_parent; N[np].height = N[best].height + 1;
        if (best != root) { if (!is_leaf(old_parent)) { if (N[old_parent].left == best) N[old_parent].left = np; else N[old_parent].right = np; } }
        else root = np;
        N[best].parent = np; N[leaf].parent = np;
        for (int i = np;;) {
            i = balance(i);
            if (!is_leaf(i)) {
                N[i].height = 1 + std::max(N[N[i].left].height, N[N[i].right].height);
                if (!refit(i)) return -1;
                if (i == root) break;
            }
            i = N[i].parent;
        }
        return leaf;
    }
    // BVH::remove(leaf)
    bool remove(int leaf) {
        const int parent = N[leaf].parent;
        drop(leaf);
        if (leaf == root) { root = 0; return true; }
        if (is_leaf(parent)) return true;
        const int sibling = N[parent].left == leaf ? N[parent].right : N[parent].left;
        if (root != parent) {
            const int gp = N[parent].parent;
            if (!is_leaf(gp)) { if (N[gp].left == parent) N[gp].left = sibling; else N[gp].right = sibling; }
            N[sibling].parent = gp;
            drop(parent);
            for (int i = gp;;) {
                i = balance(i);
                if (is_leaf(i)) continue;   // (the reference spins here too; it cannot happen: i is an ancestor)
                if (!refit(i)) return false;
                N[i].height = 1 + std::max(N[N[i].left].height, N[N[i].right].height);
                if (root == i) break;
                i = N[i].parent;
            }
        } else {
            root = sibling;
            drop(parent);
        }
        return true;
    }
    // BVH::query: callback(value) for every leaf whose box overlaps (closed test, collision.rs:22-29), right child first
    template <class F>
    void query(V3 qc, V3 qr, F&& callback) const {
        if (empty()) return;
        int stack[128]; int sp = 0;
        stack[sp++] = root;
        while (sp) {
            const CompNode& n = N[stack[--sp]];
            if (!(fabsf(qc.x - n.c.x) <= (qr.x + n.r.x) && fabsf(qc.y - n.c.y) <= (qr.y + n.r.y) && fabsf(qc.z - n.c.z) <= (qr.z + n.r.z))) continue;
            if (n.left < 0) callback(n.right);
            else if (sp + 2 <= 128) { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
    }
};

}  // namespace mgfb
