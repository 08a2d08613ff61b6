// java_hashmap_order.hpp — iteration order of a java.util.HashMap<String, ?> that only ever sees put() of NEW keys
// (what consumersPerTopic builds, LagBasedPartitionAssignor.java:412-424, and :177 iterates), INCLUDING treeified bins.
//
// Restated from the OpenJDK java.util.HashMap source (8 .. 21: the relevant methods are unchanged): putVal, resize,
// treeifyBin, TreeNode.treeify / putTreeVal / split / untreeify / balanceInsertion / rotateLeft / rotateRight /
// moveRootToFront. No JVM exists in this image, so this restatement is checked against an independently written
// Python one (tests/test_host_logic_cpu.py) and against the plain-chain model wherever no bin is treeified; it is
// NOT checked against a running JVM.
//
// What decides the iteration order: buckets are visited by index; inside a bucket the nodes' `next` links. A plain
// bin appends new keys at the tail; a resize splits a bin into a low and a high list that keep the relative order.
// The 9th key of a bin turns it into a red-black tree when the table has at least 64 slots (else the table is
// resized). A tree bin keeps a doubly linked list next to the tree: treeify() links the nodes in their old order and
// then moves the tree's ROOT to the front; a later key is linked right AFTER ITS TREE PARENT (not at the tail) and
// the (possibly new) root is moved to the front again; split() builds low / high lists in list order and
// re-treeifies a list longer than 6 (again: root to the front) or turns it back into a plain bin.
// Included at the end of java_compat.hpp (uses its utf16_units()).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace lagassign_host {

class JavaHashMapOrder {
 public:
    JavaHashMapOrder() { table_.assign(16, -1); }

    // returns the key's index (insertion number); inserts it if new
    int put_if_absent(const std::string& key) {
        auto it = index_.find(key);
        if (it != index_.end()) return it->second;
        const std::vector<uint16_t> key16 = utf16_units(key);
        uint32_t hash_code = 0;                                   // java.lang.String.hashCode
        for (uint16_t u : key16) hash_code = 31u * hash_code + u;
        const int id = (int)nodes_.size();
        Node nd;
        nd.hash = (int32_t)(hash_code ^ (hash_code >> 16));
        nodes_.push_back(nd);
        keys_.push_back(key);
        keys16_.push_back(key16);
        index_.emplace(key, id);
        put_node(id);
        if (++size_ > threshold()) resize();
        return id;
    }

    std::vector<int> order() const {
        std::vector<int> o;
        o.reserve(nodes_.size());
        for (int head : table_)
            for (int e = head; e >= 0; e = nodes_[e].next) o.push_back(e);
        return o;
    }
    const std::string& key(int id) const { return keys_[id]; }
    size_t size() const { return keys_.size(); }
    bool ever_treeified() const { return treeified_; }

 private:
    struct Node {
        int32_t hash = 0;
        int next = -1;
        // tree part (meaningful while `tree` is set)
        bool tree = false, red = false;
        int parent = -1, left = -1, right = -1, prev = -1;
    };
    static constexpr int TREEIFY_THRESHOLD = 8, UNTREEIFY_THRESHOLD = 6, MIN_TREEIFY_CAPACITY = 64;

    size_t threshold() const { return table_.size() * 3 / 4; }
    int cap() const { return (int)table_.size(); }

    // String.compareTo on the UTF-16 code units
    int compare_keys(int a, int b) const {
        const std::vector<uint16_t>& x = keys16_[a];
        const std::vector<uint16_t>& y = keys16_[b];
        const size_t n = x.size() < y.size() ? x.size() : y.size();
        for (size_t i = 0; i < n; ++i)
            if (x[i] != y[i]) return (int)x[i] - (int)y[i];
        return (int)x.size() - (int)y.size();
    }
    // direction of key x relative to tree node p, as treeify() / putTreeVal() decide it for String keys:
    // hash first (signed int compare), then compareTo; distinct strings never compare equal
    int dir_of(int x, int p) const {
        const int32_t h = nodes_[x].hash, ph = nodes_[p].hash;
        if (ph > h) return -1;
        if (ph < h) return 1;
        return compare_keys(x, p) <= 0 ? -1 : 1;
    }

    // putVal for a key that is not in the map
    void put_node(int id) {
        const int i = nodes_[id].hash & (cap() - 1);
        int p = table_[i];
        if (p < 0) { table_[i] = id; return; }
        if (nodes_[p].tree) { put_tree_val(i, id); return; }
        for (int bin_count = 0;; ++bin_count) {
            const int e = nodes_[p].next;
            if (e < 0) {
                nodes_[p].next = id;
                if (bin_count >= TREEIFY_THRESHOLD - 1) treeify_bin(i);
                return;
            }
            p = e;
        }
    }

    void treeify_bin(int index) {
        if (cap() < MIN_TREEIFY_CAPACITY) { resize(); return; }
        // replacementTreeNode for every node, same order, prev links
        int prev = -1;
        for (int e = table_[index]; e >= 0; e = nodes_[e].next) {
            Node& n = nodes_[e];
            n.tree = true; n.red = false; n.parent = n.left = n.right = -1;
            n.prev = prev;
            prev = e;
        }
        treeify(table_[index]);
    }

    // TreeNode.treeify: forms the tree of the nodes linked from `head`, then moves its root to the front of the bin
    void treeify(int head) {
        treeified_ = true;
        int root = -1;
        for (int x = head, next; x >= 0; x = next) {
            next = nodes_[x].next;
            nodes_[x].left = nodes_[x].right = -1;
            if (root < 0) {
                nodes_[x].parent = -1;
                nodes_[x].red = false;
                root = x;
            } else {
                for (int p = root;;) {
                    const int dir = dir_of(x, p);
                    const int xp = p;
                    p = (dir <= 0) ? nodes_[p].left : nodes_[p].right;
                    if (p < 0) {
                        nodes_[x].parent = xp;
                        if (dir <= 0) nodes_[xp].left = x; else nodes_[xp].right = x;
                        root = balance_insertion(root, x);
                        break;
                    }
                }
            }
        }
        move_root_to_front(root);
    }

    // TreeNode.putTreeVal for a new key: linked after its tree parent
    void put_tree_val(int index, int x) {
        int root = table_[index];
        while (nodes_[root].parent >= 0) root = nodes_[root].parent;       // root()
        Node& nx = nodes_[x];
        nx.tree = true; nx.left = nx.right = -1;
        for (int p = root;;) {
            const int dir = dir_of(x, p);
            const int xp = p;
            p = (dir <= 0) ? nodes_[p].left : nodes_[p].right;
            if (p < 0) {
                const int xpn = nodes_[xp].next;
                nodes_[x].next = xpn;
                if (dir <= 0) nodes_[xp].left = x; else nodes_[xp].right = x;
                nodes_[xp].next = x;
                nodes_[x].parent = nodes_[x].prev = xp;
                if (xpn >= 0) nodes_[xpn].prev = x;
                move_root_to_front(balance_insertion(root, x));
                return;
            }
        }
    }

    void move_root_to_front(int root) {
        if (root < 0) return;
        const int index = nodes_[root].hash & (cap() - 1);
        const int first = table_[index];
        if (root != first) {
            table_[index] = root;
            const int rp = nodes_[root].prev, rn = nodes_[root].next;
            if (rn >= 0) nodes_[rn].prev = rp;
            if (rp >= 0) nodes_[rp].next = rn;
            if (first >= 0) nodes_[first].prev = root;
            nodes_[root].next = first;
            nodes_[root].prev = -1;
        }
    }

    int rotate_left(int root, int p) {
        int r, pp, rl;
        if (p >= 0 && (r = nodes_[p].right) >= 0) {
            if ((rl = nodes_[p].right = nodes_[r].left) >= 0) nodes_[rl].parent = p;
            if ((pp = nodes_[r].parent = nodes_[p].parent) < 0) { root = r; nodes_[r].red = false; }
            else if (nodes_[pp].left == p) nodes_[pp].left = r;
            else nodes_[pp].right = r;
            nodes_[r].left = p;
            nodes_[p].parent = r;
        }
        return root;
    }
    int rotate_right(int root, int p) {
        int l, pp, lr;
        if (p >= 0 && (l = nodes_[p].left) >= 0) {
            if ((lr = nodes_[p].left = nodes_[l].right) >= 0) nodes_[lr].parent = p;
            if ((pp = nodes_[l].parent = nodes_[p].parent) < 0) { root = l; nodes_[l].red = false; }
            else if (nodes_[pp].right == p) nodes_[pp].right = l;
            else nodes_[pp].left = l;
            nodes_[l].right = p;
            nodes_[p].parent = l;
        }
        return root;
    }
    int balance_insertion(int root, int x) {
        nodes_[x].red = true;
        for (int xp, xpp, xppl, xppr;;) {
            if ((xp = nodes_[x].parent) < 0) { nodes_[x].red = false; return x; }
            else if (!nodes_[xp].red || (xpp = nodes_[xp].parent) < 0) return root;
            if (xp == (xppl = nodes_[xpp].left)) {
                if ((xppr = nodes_[xpp].right) >= 0 && nodes_[xppr].red) {
                    nodes_[xppr].red = false; nodes_[xp].red = false; nodes_[xpp].red = true;
                    x = xpp;
                } else {
                    if (x == nodes_[xp].right) {
                        root = rotate_left(root, x = xp);
                        xpp = (xp = nodes_[x].parent) < 0 ? -1 : nodes_[xp].parent;
                    }
                    if (xp >= 0) {
                        nodes_[xp].red = false;
                        if (xpp >= 0) { nodes_[xpp].red = true; root = rotate_right(root, xpp); }
                    }
                }
            } else {
                if (xppl >= 0 && nodes_[xppl].red) {
                    nodes_[xppl].red = false; nodes_[xp].red = false; nodes_[xpp].red = true;
                    x = xpp;
                } else {
                    if (x == nodes_[xp].left) {
                        root = rotate_right(root, x = xp);
                        xpp = (xp = nodes_[x].parent) < 0 ? -1 : nodes_[xp].parent;
                    }
                    if (xp >= 0) {
                        nodes_[xp].red = false;
                        if (xpp >= 0) { nodes_[xpp].red = true; root = rotate_left(root, xpp); }
                    }
                }
            }
        }
    }

    // HashMap.resize: every bin splits into a low list (same index) and a high list (index + old capacity) that keep
    // the order of the bin's `next` links; tree bins then untreeify (<= 6 nodes) or re-treeify (when both lists exist)
    void resize() {
        const int old_cap = cap();
        std::vector<int> old;
        old.swap(table_);
        table_.assign((size_t)old_cap * 2, -1);
        for (int j = 0; j < old_cap; ++j) {
            int e = old[j];
            if (e < 0) continue;
            const bool was_tree = nodes_[e].tree;
            int lo_head = -1, lo_tail = -1, hi_head = -1, hi_tail = -1, lc = 0, hc = 0;
            for (int next; e >= 0; e = next) {
                next = nodes_[e].next;
                nodes_[e].next = -1;
                if ((nodes_[e].hash & old_cap) == 0) {
                    nodes_[e].prev = lo_tail;
                    if (lo_tail < 0) lo_head = e; else nodes_[lo_tail].next = e;
                    lo_tail = e; ++lc;
                } else {
                    nodes_[e].prev = hi_tail;
                    if (hi_tail < 0) hi_head = e; else nodes_[hi_tail].next = e;
                    hi_tail = e; ++hc;
                }
            }
            if (!was_tree) {
                table_[j] = lo_head;
                table_[j + old_cap] = hi_head;
                continue;
            }
            if (lo_head >= 0) {
                if (lc <= UNTREEIFY_THRESHOLD) { untreeify(lo_head); table_[j] = lo_head; }
                else { table_[j] = lo_head; if (hi_head >= 0) treeify(lo_head); }      // (else it is the old tree, unchanged)
            }
            if (hi_head >= 0) {
                if (hc <= UNTREEIFY_THRESHOLD) { untreeify(hi_head); table_[j + old_cap] = hi_head; }
                else { table_[j + old_cap] = hi_head; if (lo_head >= 0) treeify(hi_head); }
            }
        }
    }
    void untreeify(int head) {
        for (int e = head; e >= 0; e = nodes_[e].next) {
            Node& n = nodes_[e];
            n.tree = false; n.red = false; n.parent = n.left = n.right = n.prev = -1;
        }
    }

    std::vector<int> table_;
    std::vector<Node> nodes_;
    std::vector<std::string> keys_;
    std::vector<std::vector<uint16_t>> keys16_;
    std::unordered_map<std::string, int> index_;
    size_t size_ = 0;
    bool treeified_ = false;
};

}  // namespace lagassign_host
