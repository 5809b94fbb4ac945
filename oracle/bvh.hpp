// ORACLE (test infrastructure, NOT product code).  Restates /root/reference/src/pool.rs
// (free-list slab) and src/bvh.rs (incremental AABB tree: insert / remove / balance / query)
// operation-for-operation, for B = AABB.  Citations are file:line of the reference.
#pragma once
#include <stdexcept>
#include <vector>
#include "collision.hpp"

namespace mgfo {

// pool.rs:25-113
template <class T>
struct Pool {
    enum Tag { FreeListEnd, FreeListPtr, Occupied };
    struct Entry { Tag tag; size_t next_free; T item; };
    size_t len = 0;
    bool has_free = false;
    size_t free_list = 0;
    std::vector<Entry> entries;

    bool empty() const { return len == 0; }
    void clear() { len = 0; has_free = false; entries.clear(); }
    size_t push(const T& item) {  // pool.rs:81-96
        len += 1;
        if (has_free) {
            size_t free_item = free_list;
            Entry& e = entries[free_item];
            if (e.tag == FreeListEnd) has_free = false;
            else if (e.tag == FreeListPtr) free_list = e.next_free;
            else throw std::logic_error("unreachable");
            e.tag = Occupied; e.item = item;
            return free_item;
        }
        size_t i = entries.size();
        entries.push_back(Entry{Occupied, 0, item});
        return i;
    }
    T remove(size_t i) {  // pool.rs:100-113
        Entry& e = entries[i];
        if (e.tag != Occupied) throw std::out_of_range("index is not occupied");
        T item = e.item;
        if (has_free) { e.tag = FreeListPtr; e.next_free = free_list; }
        else { e.tag = FreeListEnd; }
        has_free = true; free_list = i;
        len -= 1;
        return item;
    }
    T& operator[](size_t i) {
        if (i >= entries.size() || entries[i].tag != Occupied) throw std::out_of_range("index is not occupied");
        return entries[i].item;
    }
    const T& operator[](size_t i) const {
        if (i >= entries.size() || entries[i].tag != Occupied) throw std::out_of_range("index is not occupied");
        return entries[i].item;
    }
};

// bvh.rs:29-44
template <class V>
struct BVH {
    struct Node {
        int height; size_t parent; AABB bounds;
        bool is_leaf; V val; size_t c1, c2;
    };
    size_t root_ = 0;
    Pool<Node> pool;

    bool empty() const { return pool.empty(); }
    void clear() { root_ = 0; pool.clear(); }
    size_t insert_node(const AABB& b, bool leaf, const V& val, size_t c1, size_t c2) {  // bvh.rs:114-121
        return pool.push(Node{-1, 0, b, leaf, val, c1, c2});
    }
    // bvh.rs:125-217
    size_t insert(const AABB& bounds, const V& val) {
        size_t leaf = insert_node(bounds, true, val, 0, 0);
        if (pool.len == 1) { root_ = leaf; return leaf; }
        size_t best = root_;
        for (;;) {
            if (!pool[best].is_leaf) {
                size_t child1 = pool[best].c1, child2 = pool[best].c2;
                AABB curr_bounds = pool[best].bounds;
                float area = surface_area(curr_bounds);
                AABB combined_bounds = aabb_combine(curr_bounds, bounds);
                float combined_area = surface_area(combined_bounds);
                float no_descent_cost = combined_area * 2.0f;
                float inheritance_cost = (combined_area - area) * 2.0f;
                auto child_cost = [&](size_t child) -> float {
                    if (!pool[child].is_leaf) {
                        float old_area = surface_area(pool[child].bounds);
                        float new_area = surface_area(aabb_combine(bounds, pool[child].bounds));
                        return new_area - old_area + inheritance_cost;
                    } else {
                        return surface_area(aabb_combine(bounds, pool[child].bounds)) + inheritance_cost;
                    }
                };
                float child1_cost = child_cost(child1);
                float child2_cost = child_cost(child2);
                if (no_descent_cost < child1_cost && no_descent_cost < child2_cost) break;
                best = child1_cost < child2_cost ? child1 : child2;
            } else {
                break;
            }
        }
        size_t old_parent = pool[best].parent;
        AABB best_bounds = pool[best].bounds;
        size_t new_parent = insert_node(aabb_combine(bounds, best_bounds), false, V(), best, leaf);
        pool[new_parent].parent = old_parent;
        pool[new_parent].height = pool[best].height + 1;
        if (best != root_) {
            Node& op = pool[old_parent];
            if (!op.is_leaf) {
                if (op.c1 == best) op.c1 = new_parent; else op.c2 = new_parent;
            }
        } else {
            root_ = new_parent;
        }
        pool[best].parent = new_parent;
        pool[leaf].parent = new_parent;
        size_t i = pool[leaf].parent;
        for (;;) {
            i = balance(i);
            if (!pool[i].is_leaf) {
                size_t child1 = pool[i].c1, child2 = pool[i].c2;
                pool[i].height = 1 + std::max(pool[child1].height, pool[child2].height);
                pool[i].bounds = aabb_combine(pool[child1].bounds, pool[child2].bounds);
                if (i == root_) break;
            }
            i = pool[i].parent;
        }
        return leaf;
    }
    // bvh.rs:220-260
    void remove(size_t leaf) {
        size_t parent = pool[leaf].parent;
        pool.remove(leaf);
        if (leaf == root_) { root_ = 0; return; }
        if (!pool[parent].is_leaf) {
            size_t child1 = pool[parent].c1, child2 = pool[parent].c2;
            size_t sibling = child1 == leaf ? child2 : child1;
            if (root_ != parent) {
                size_t grand_parent = pool[parent].parent;
                Node& gp = pool[grand_parent];
                if (!gp.is_leaf) {
                    if (gp.c1 == parent) gp.c1 = sibling; else gp.c2 = sibling;
                }
                pool[sibling].parent = grand_parent;
                pool.remove(parent);
                size_t i = grand_parent;
                for (;;) {
                    i = balance(i);
                    if (!pool[i].is_leaf) {
                        size_t c1 = pool[i].c1, c2 = pool[i].c2;
                        pool[i].bounds = aabb_combine(pool[c1].bounds, pool[c2].bounds);
                        pool[i].height = 1 + std::max(pool[c1].height, pool[c2].height);
                        if (root_ == i) break;
                        i = pool[i].parent;
                    }
                }
            } else {
                root_ = sibling;
                pool.remove(parent);
            }
        }
    }
    size_t root() const {
        if (empty()) throw std::runtime_error("BVH is empty, there is no root node");
        return root_;
    }
    const AABB& operator[](size_t i) const { return pool[i].bounds; }  // bvh.rs:483-492
    // bvh.rs:283-310: stack DFS, pushes lchild then rchild => right child visited first.
    template <class F>
    void query(const AABB& arg_bounds, F&& callback) const {
        if (empty()) return;
        std::vector<size_t> stack;
        stack.push_back(root_);
        while (!stack.empty()) {
            size_t top = stack.back(); stack.pop_back();
            const Node& n = pool[top];
            if (overlaps(arg_bounds, n.bounds)) {
                if (n.is_leaf) callback(n.val);
                else { stack.push_back(n.c1); stack.push_back(n.c2); }
            }
        }
    }
    // bvh.rs:345-369
    template <class F>
    void raytrace(const Ray& ray, F&& callback, float DT = INF) const {
        if (empty()) return;
        std::vector<size_t> stack;
        stack.push_back(root_);
        while (!stack.empty()) {
            size_t top = stack.back(); stack.pop_back();
            const Node& n = pool[top];
            Intersection inter;
            if (intersection(ray, n.bounds, &inter, DT)) {
                if (n.is_leaf) callback(n.val, inter);
                else { stack.push_back(n.c1); stack.push_back(n.c2); }
            }
        }
    }
    // bvh.rs:371-480
    size_t balance(size_t a) {
        if (pool[a].height < 2) return a;
        if (!pool[a].is_leaf) {
            size_t b = pool[a].c1, c = pool[a].c2;
            if (pool[c].height > pool[b].height + 1) {
                if (!pool[c].is_leaf) {
                    size_t f = pool[c].c1, g = pool[c].c2;
                    pool[c].parent = pool[a].parent;
                    pool[a].parent = c;
                    if (root_ == a) {
                        root_ = c;
                    } else if (!pool[pool[c].parent].is_leaf) {
                        size_t parent = pool[c].parent;
                        if (pool[parent].c1 == a) pool[parent].c1 = c; else pool[parent].c2 = c;
                    }
                    if (pool[f].height > pool[g].height) {
                        pool[c].c1 = a; pool[c].c2 = f;
                        pool[a].c1 = b; pool[a].c2 = g;
                        pool[g].parent = a;
                        pool[a].bounds = aabb_combine(pool[b].bounds, pool[g].bounds);
                        pool[c].bounds = aabb_combine(pool[a].bounds, pool[f].bounds);
                        pool[a].height = 1 + std::max(pool[b].height, pool[g].height);
                        pool[c].height = 1 + std::max(pool[a].height, pool[f].height);
                    } else {
                        pool[c].c1 = a; pool[c].c2 = g;
                        pool[a].c1 = b; pool[a].c2 = f;
                        pool[f].parent = a;
                        pool[a].bounds = aabb_combine(pool[b].bounds, pool[f].bounds);
                        pool[c].bounds = aabb_combine(pool[a].bounds, pool[g].bounds);
                        pool[a].height = 1 + std::max(pool[b].height, pool[f].height);
                        pool[c].height = 1 + std::max(pool[a].height, pool[g].height);
                    }
                }
                return c;
            }
            if (pool[b].height > pool[c].height + 1) {
                if (!pool[b].is_leaf) {
                    size_t d = pool[b].c1, e = pool[b].c2;
                    pool[b].parent = pool[a].parent;
                    pool[a].parent = b;
                    if (root_ == a) {
                        root_ = b;
                    } else if (!pool[pool[b].parent].is_leaf) {
                        size_t parent = pool[b].parent;
                        if (pool[parent].c1 == a) pool[parent].c1 = b; else pool[parent].c2 = b;
                    }
                    if (pool[d].height > pool[e].height) {
                        pool[b].c1 = a; pool[b].c2 = d;
                        pool[a].c1 = e; pool[a].c2 = c;
                        pool[e].parent = a;
                        pool[a].bounds = aabb_combine(pool[c].bounds, pool[e].bounds);
                        pool[b].bounds = aabb_combine(pool[a].bounds, pool[d].bounds);
                        pool[a].height = 1 + std::max(pool[c].height, pool[e].height);
                        pool[b].height = 1 + std::max(pool[a].height, pool[d].height);
                    } else {
                        pool[b].c1 = a; pool[b].c2 = e;
                        pool[a].c1 = d; pool[a].c2 = c;
                        pool[d].parent = a;
                        pool[a].bounds = aabb_combine(pool[c].bounds, pool[d].bounds);
                        pool[b].bounds = aabb_combine(pool[a].bounds, pool[e].bounds);
                        pool[a].height = 1 + std::max(pool[c].height, pool[d].height);
                        pool[b].height = 1 + std::max(pool[a].height, pool[e].height);
                    }
                }
                return b;
            }
        }
        return a;
    }
};

}  // namespace mgfo
