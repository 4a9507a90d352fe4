// ccu_layout.h - commit-time layout builders (host C++): the value-carrying octree layout, the march ("air") layout and
// the BVH stage layout.  Included by chunkycu.cu only.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ccu_device.cuh"
#include "ccu_march.cuh"


// ------------------------------------------------------------------------------------------------------
// commit-time traversal layout (see DScene::top / DScene::wide)
// ------------------------------------------------------------------------------------------------------
namespace {
struct WideLayout {
    std::vector<unsigned> top, wide;
    int cell_level = 0, top_log2 = 0;
    bool ok = true;
};

inline unsigned wide_leaf(int word, int level, bool &ok) {
    long long v = -(long long)word;
    unsigned val;
    if (v == 0x7FFFFFFELL) val = CCU_WIDE_ANY;
    else if (v >= 0 && v < (long long)CCU_WIDE_ANY) val = (unsigned)v;
    else { ok = false; val = 0; }
    return CCU_WIDE_LEAF | ((unsigned)level << 26) | val;
}

unsigned wide_node(const int *tree, size_t n, int word, int lvl, WideLayout &b) {
    if (lvl < 2 || (size_t)word + 7 >= n) { b.ok = false; return wide_leaf(0, 0, b.ok); }
    const size_t idx = b.wide.size() / 64;
    if (idx >= 0x7FFFFFFFu) { b.ok = false; return wide_leaf(0, 0, b.ok); }
    b.wide.resize(b.wide.size() + 64);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                unsigned e;
                int c1 = tree[(size_t)word + ((((i >> 1) & 1) << 2) | (((j >> 1) & 1) << 1) | ((k >> 1) & 1))];
                if (c1 <= 0) {
                    e = wide_leaf(c1, lvl - 1, b.ok);
                } else if ((size_t)c1 + 7 >= n) {
                    b.ok = false;
                    e = wide_leaf(0, 0, b.ok);
                } else {
                    int c2 = tree[(size_t)c1 + (((i & 1) << 2) | ((j & 1) << 1) | (k & 1))];
                    e = c2 <= 0 ? wide_leaf(c2, lvl - 2, b.ok) : wide_node(tree, n, c2, lvl - 2, b);
                }
                b.wide[idx * 64 + ((i << 4) | (j << 2) | k)] = e;
            }
    return (unsigned)idx;
}

// "Air layout" for the march loop (ccu_march.cuh): top table over cells of level `cell_level` (>= 4), 64-ary nodes down to
// level 4 for deep worlds, and one 1-KiB brick of 2-bit voxel codes per 16^3 cell that is not a single leaf.  The march loop
// needs nothing else (octree.h:89-106: an air leaf is left through its cube, anything else goes to the block test); a leaf
// lookup is one table load plus at most one brick load.
struct AirLayout {
    std::vector<unsigned> top, wide, bricks;
    std::vector<unsigned char> wide_level;     // level of every 64-ary node (its entries sit two levels below)
    int cell_level = 4, top_log2 = 0;
    bool ok = true;
};

// The top table is filled slab by slab (x ranges) on several host threads; every thread appends to vectors of its own, and the
// indices it handed out are shifted by the sizes of the slabs before it when the parts are joined.
inline int layout_threads(int dim) {
    if (dim < 8) return 1;
    if (const char *e = getenv("CCU_COMMIT_THREADS")) return std::max(1, std::min(dim, atoi(e)));
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min({hw ? hw : 4u, (unsigned)dim, 16u}));
}

inline unsigned air_leaf(int word, int level) { return CCU_WIDE_LEAF | ((word == 0 ? (unsigned)level : 31u) << 26); }
inline int oct_child(int x, int y, int z) { return ((x & 1) << 2) | ((y & 1) << 1) | (z & 1); }

// brick of the 16^3 cell rooted at the branch node `word` (level 4)
unsigned air_brick(const int *tree, size_t n, int word, AirLayout &b) {
    const size_t idx = b.bricks.size() / 256;
    if (idx >= 0x7FFFFFu || (size_t)word + 7 >= n) { b.ok = false; return air_leaf(1, 0); }
    b.bricks.resize(b.bricks.size() + 256, 0u);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                unsigned *blk = &b.bricks[idx * 256 + (size_t)((i << 4) | (j << 2) | k) * 4];
                unsigned uniform = 0;
                bool mixed = false;
                int c2 = 0;
                const int c1 = tree[(size_t)word + oct_child(i >> 1, j >> 1, k >> 1)];        // level 3
                if (c1 <= 0) {
                    uniform = c1 == 0 ? (CCU_BRICK_UNIFORM | 3u) : 0u;
                } else if ((size_t)c1 + 7 >= n) {
                    b.ok = false;
                } else {
                    c2 = tree[(size_t)c1 + oct_child(i, j, k)];                               // level 2
                    if (c2 <= 0) uniform = c2 == 0 ? (CCU_BRICK_UNIFORM | 2u) : 0u;
                    else if ((size_t)c2 + 7 >= n) b.ok = false;
                    else mixed = true;
                }
                if (!mixed) {
                    blk[0] = blk[1] = blk[2] = blk[3] = uniform;
                    continue;
                }
                for (int x = 0; x < 4; x++)
                    for (int y = 0; y < 4; y++)
                        for (int z = 0; z < 4; z++) {
                            unsigned code;
                            const int d1 = tree[(size_t)c2 + oct_child(x >> 1, y >> 1, z >> 1)];   // level 1
                            if (d1 <= 0) {
                                code = d1 == 0 ? 2u : 0u;
                            } else if ((size_t)d1 + 7 >= n) {
                                b.ok = false;
                                code = 0;
                            } else {
                                const int d2 = tree[(size_t)d1 + oct_child(x, y, z)];              // level 0
                                if (d2 > 0) b.ok = false;                                           // deeper than the declared depth
                                code = d2 == 0 ? 1u : 0u;
                            }
                            blk[x] |= code << ((((y << 2) | z)) * 2);
                        }
            }
    return (unsigned)idx;
}

unsigned air_node(const int *tree, size_t n, int word, int lvl, AirLayout &b) {
    if (lvl == 4) return air_brick(tree, n, word, b);
    if (lvl < 6 || (size_t)word + 7 >= n) { b.ok = false; return air_leaf(1, 0); }
    const size_t idx = b.wide.size() / 64;
    if (idx >= 0x1FFFFFFu) { b.ok = false; return air_leaf(1, 0); }
    b.wide.resize(b.wide.size() + 64);
    b.wide_level.push_back((unsigned char)lvl);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 4; k++) {
                unsigned e;
                const int c1 = tree[(size_t)word + oct_child(i >> 1, j >> 1, k >> 1)];
                if (c1 <= 0) {
                    e = air_leaf(c1, lvl - 1);
                } else if ((size_t)c1 + 7 >= n) {
                    b.ok = false;
                    e = air_leaf(1, 0);
                } else {
                    const int c2 = tree[(size_t)c1 + oct_child(i, j, k)];
                    e = c2 <= 0 ? air_leaf(c2, lvl - 2) : air_node(tree, n, c2, lvl - 2, b);
                }
                b.wide[idx * 64 + ((i << 4) | (j << 2) | k)] = e;
            }
    return (unsigned)idx;
}

// Octrees shallower than a brick (depth < 4): one brick whose corner [0, 2^depth)^3 holds the tree, filled voxel by voxel
// from the root descent (octree.h:81-88); the march loop never looks outside the cube (bounds check octree.h:76-78).
void air_small_brick(const int *tree, size_t n, int depth, AirLayout &b) {
    b.bricks.assign(256, 0u);
    const int size = 1 << depth;
    for (int x = 0; x < size; x++)
        for (int y = 0; y < size; y++)
            for (int z = 0; z < size; z++) {
                int level = depth, word = tree[0];
                while (word > 0 && level > 0) {
                    level--;
                    const size_t at = (size_t)word + oct_child(x >> level, y >> level, z >> level);
                    if (at >= n) { b.ok = false; word = -1; break; }
                    word = tree[at];
                }
                if (word > 0) { b.ok = false; word = -1; }
                unsigned *blk = &b.bricks[(size_t)((((x >> 2) & 3) << 4) | (((y >> 2) & 3) << 2) | ((z >> 2) & 3)) * 4];
                if (word == 0 && level >= 2) {
                    blk[x & 3] = CCU_BRICK_UNIFORM | (unsigned)level;
                } else {
                    const unsigned code = word == 0 ? (unsigned)level + 1u : 0u;
                    blk[x & 3] |= code << (((((y & 3) << 2) | (z & 3))) * 2);
                }
            }
}

// cells [x0, x1) of the top table, into `b` (indices local to b)
inline void air_fill_slab(const int *tree, size_t n, int depth, int cl, int dim, int x0, int x1, unsigned *top, AirLayout &b) {
    for (int x = x0; x < x1 && b.ok; x++)
        for (int y = 0; y < dim; y++)
            for (int z = 0; z < dim; z++) {
                int level = depth;
                int word = tree[0];
                while (word > 0 && level > cl) {
                    level--;
                    const int sh = level - cl;
                    const size_t at = (size_t)word + oct_child(x >> sh, y >> sh, z >> sh);
                    if (at >= n) { b.ok = false; word = 0; break; }
                    word = tree[at];
                }
                top[((size_t)x * dim + y) * dim + z] = word <= 0 ? air_leaf(word, level) : air_node(tree, n, word, cl, b);
            }
}

AirLayout build_air_layout(const int *tree, size_t n, int depth) {
    AirLayout b;
    if (depth < 4) {
        b.cell_level = 4;
        b.top_log2 = 0;
        if (tree[0] <= 0) b.top.assign(1, air_leaf(tree[0], depth));
        else { b.top.assign(1, 0u); air_small_brick(tree, n, depth, b); }
    } else {
        int cl = std::max(depth - 7, 4);             // top table of at most 128^3 cells
        if (cl & 1) cl++;
        if (cl > depth) cl = depth & ~1;
        b.cell_level = cl;
        b.top_log2 = depth - cl;
        const int dim = 1 << b.top_log2;
        b.top.assign((size_t)dim * dim * dim, 0u);
        const int nt = layout_threads(dim);
        if (nt == 1) {
            air_fill_slab(tree, n, depth, cl, dim, 0, dim, b.top.data(), b);
        } else {
            std::vector<AirLayout> part(nt);
            std::vector<std::thread> th;
            for (int t = 0; t < nt; t++)
                th.emplace_back([&, t] { air_fill_slab(tree, n, depth, cl, dim, dim * t / nt, dim * (t + 1) / nt, b.top.data(), part[t]); });
            for (auto &t : th) t.join();
            for (int t = 0; t < nt; t++) {
                const unsigned wide_off = (unsigned)(b.wide.size() / 64), brick_off = (unsigned)(b.bricks.size() / 256);
                AirLayout &p = part[t];
                b.ok = b.ok && p.ok;
                // entries below the leaf bit are indices: of bricks where they sit at level 4, of 64-ary nodes above that
                for (size_t j = 0; j < p.wide_level.size(); j++) {
                    const unsigned off = p.wide_level[j] - 2 == 4 ? brick_off : wide_off;
                    for (int e = 0; e < 64; e++)
                        if (!(p.wide[j * 64 + e] & CCU_WIDE_LEAF)) p.wide[j * 64 + e] += off;
                }
                const unsigned top_off = cl == 4 ? brick_off : wide_off;
                const size_t c0 = (size_t)(dim * t / nt) * dim * dim, c1 = (size_t)(dim * (t + 1) / nt) * dim * dim;
                for (size_t c = c0; c < c1; c++)
                    if (!(b.top[c] & CCU_WIDE_LEAF)) b.top[c] += top_off;
                b.wide.insert(b.wide.end(), p.wide.begin(), p.wide.end());
                b.wide_level.insert(b.wide_level.end(), p.wide_level.begin(), p.wide_level.end());
                b.bricks.insert(b.bricks.end(), p.bricks.begin(), p.bricks.end());
            }
            if (b.wide.size() / 64 >= 0x1FFFFFFu || b.bricks.size() / 256 >= 0x7FFFFFu) b.ok = false;
        }
    }
    if (b.wide.empty()) b.wide.assign(64, air_leaf(1, 0));
    if (b.bricks.empty()) b.bricks.assign(256, 0u);
    return b;
}

// Traversal layout of a packed BVH (PackedBvhNode.java:22-31: 7 ints per node, first child at node + 7, second child at
// node[0]) for the BVH stage of ccu_queue.cuh: one 64-byte record per inner node holding BOTH children's boxes
// (bvh.h:73-91 fetches exactly those at every inner node) and a reference per child, and 16-byte aligned triangle
// blocks.  ref >= 0: record index; ref < 0: leaf, -(1 + offset of its block in `tris`, in units of 8 words).
struct BvhLayout {
    std::vector<int> rec;
    int root = 0;
    bool ok = true;
};
struct TriRepack {
    std::vector<int> tris;                       // per leaf: {count, 0 x 7} + count x (20 words (PackedTriangle.java:46-78) + 4 pad)
    int add(const std::vector<int> &trigs, int prim, bool &ok) {
        if (prim < 0 || (size_t)prim >= trigs.size()) { ok = false; return 0; }
        const int count = trigs[(size_t)prim];
        if (count < 0 || (size_t)prim + 1 + (size_t)count * 20 > trigs.size()) { ok = false; return 0; }
        const int off = (int)(tris.size() / 8);      // in units of 32 bytes
        tris.push_back(count);
        tris.insert(tris.end(), 7, 0);
        for (int i = 0; i < count; i++) {
            tris.insert(tris.end(), trigs.begin() + prim + 1 + (size_t)i * 20, trigs.begin() + prim + 1 + (size_t)(i + 1) * 20);
            tris.insert(tris.end(), 4, 0);
        }
        return off;
    }
};

int bvh_ref(const std::vector<int> &bvh, const std::vector<int> &trigs, size_t node, int depth, BvhLayout &b, TriRepack &tr,
            std::unordered_map<int, int> &leaf_map) {
    if (!b.ok) return -1;
    // deeper than the traversal stack of the reference (int nodesToVisit[64], bvh.h:38): undefined there, refused here
    if (node + 6 >= bvh.size() || depth >= 64) { b.ok = false; return -1; }
    const int head = bvh[node];
    if (head <= 0) {
        const int prim = -head;
        auto it = leaf_map.find(prim);   // a leaf block referenced twice (both BVHs share the palette) is stored once
        int off;
        if (it != leaf_map.end()) {
            off = it->second;
        } else {
            off = tr.add(trigs, prim, b.ok);
            leaf_map.emplace(prim, off);
        }
        return -(1 + off);
    }
    const size_t left = node + 7, right = (size_t)head;
    if (left + 6 >= bvh.size() || right + 6 >= bvh.size()) { b.ok = false; return -1; }
    const size_t r = b.rec.size() / 16;
    b.rec.resize(b.rec.size() + 16, 0);
    // two 32-byte halves of the same shape: {box (6 floats), ref, 0} of the first child, then of the second child
    for (int i = 0; i < 6; i++) {
        b.rec[r * 16 + i] = bvh[left + 1 + i];
        b.rec[r * 16 + 8 + i] = bvh[right + 1 + i];
    }
    const int rl = bvh_ref(bvh, trigs, left, depth + 1, b, tr, leaf_map);
    const int rr = bvh_ref(bvh, trigs, right, depth + 1, b, tr, leaf_map);
    b.rec[r * 16 + 6] = rl;
    b.rec[r * 16 + 14] = rr;
    return (int)r;
}

inline void wide_fill_slab(const int *tree, size_t n, int depth, int cl, int dim, int x0, int x1, unsigned *top, WideLayout &b) {
    for (int x = x0; x < x1 && b.ok; x++)
        for (int y = 0; y < dim; y++)
            for (int z = 0; z < dim; z++) {
                int level = depth;
                int word = tree[0];
                while (word > 0 && level > cl) {
                    level--;
                    int sh = level - cl;
                    size_t at = (size_t)word + ((((x >> sh) & 1) << 2) | (((y >> sh) & 1) << 1) | ((z >> sh) & 1));
                    if (at >= n) { b.ok = false; word = 0; break; }
                    word = tree[at];
                }
                top[((size_t)x * dim + y) * dim + z] = word <= 0 ? wide_leaf(word, level, b.ok) : wide_node(tree, n, word, cl, b);
            }
}

WideLayout build_wide_layout(const int *tree, size_t n, int depth) {
    WideLayout b;
    int cl = std::max(depth - 7, 4);
    if (const char *e = getenv("CCU_CELL_LEVEL")) cl = std::max(0, atoi(e));   // tuning knob: level of the top table's cells
    if (cl & 1) cl++;
    if (cl > depth) cl = depth & ~1;
    b.cell_level = cl;
    b.top_log2 = depth - cl;
    const int dim = 1 << b.top_log2;
    b.top.assign((size_t)dim * dim * dim, 0u);
    const int nt = layout_threads(dim);
    if (nt == 1) {
        wide_fill_slab(tree, n, depth, cl, dim, 0, dim, b.top.data(), b);
    } else {
        std::vector<WideLayout> part(nt);
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++)
            th.emplace_back([&, t] { wide_fill_slab(tree, n, depth, cl, dim, dim * t / nt, dim * (t + 1) / nt, b.top.data(), part[t]); });
        for (auto &t : th) t.join();
        for (int t = 0; t < nt; t++) {
            const unsigned off = (unsigned)(b.wide.size() / 64);
            WideLayout &p = part[t];
            b.ok = b.ok && p.ok;
            for (unsigned &e : p.wide)
                if (!(e & CCU_WIDE_LEAF)) e += off;           // every index refers to a 64-ary node
            const size_t c0 = (size_t)(dim * t / nt) * dim * dim, c1 = (size_t)(dim * (t + 1) / nt) * dim * dim;
            for (size_t c = c0; c < c1; c++)
                if (!(b.top[c] & CCU_WIDE_LEAF)) b.top[c] += off;
            b.wide.insert(b.wide.end(), p.wide.begin(), p.wide.end());
        }
        if (b.wide.size() / 64 >= 0x7FFFFFFFu) b.ok = false;
    }
    if (b.wide.empty()) b.wide.assign(64, wide_leaf(0, 0, b.ok));
    return b;
}

// 16-byte-vectorised palettes (DScene::block_rec / mat_rec / quad_rec / aabb_rec): the reference's scalar int[] palettes
// (PackedBlock.java:79-85, PackedMaterial.java:74-100, PackedQuad.java:41-66, PackedAabb.java:75-102) re-laid out so that a
// block test reads whole records with 128-bit loads.  A block-palette position whose entry cannot be translated (model pointer
// outside its palette) keeps type 0x7FFFFFFF, which no branch of intersect_block accepts - the same miss the reference's
// out-of-range read would most likely produce, but defined.
struct PaletteRecs {
    std::vector<int> block, mat, quad, aabb;
    bool ok = true;
};

inline PaletteRecs build_palette_recs(const std::vector<int> &bp, const std::vector<int> &mp, const int *quads, size_t n_quads,
                                      const int *aabbs, size_t n_aabbs) {
    PaletteRecs r;
    const size_t n_mat = mp.size() / 6;
    r.mat.assign(std::max<size_t>(n_mat, 1) * 8, 0);
    for (size_t i = 0; i < n_mat; i++)
        for (int k = 0; k < 6; k++) r.mat[i * 8 + k] = mp[i * 6 + k];
    std::unordered_map<int, int> quad_at, aabb_at;     // model pointer -> offset in units of 16 bytes
    auto model = [&](const int *src, size_t n, int ptr, int words, std::vector<int> &dst, std::unordered_map<int, int> &at, int &out) {
        auto it = at.find(ptr);
        if (it != at.end()) { out = it->second; return true; }
        if (ptr < 0 || (size_t)ptr >= n) return false;
        const int count = src[ptr];
        if (count < 0 || (size_t)ptr + 1 + (size_t)count * words > n) return false;
        out = (int)(dst.size() / 4);
        dst.push_back(count); dst.push_back(0); dst.push_back(0); dst.push_back(0);
        for (int i = 0; i < count; i++) {
            for (int k = 0; k < 16; k++) dst.push_back(k < words ? src[(size_t)ptr + 1 + (size_t)i * words + k] : 0);
        }
        at.emplace(ptr, out);
        return true;
    };
    const size_t n_pos = bp.size() >= 2 ? bp.size() - 1 : 0;
    r.block.assign(std::max<size_t>(n_pos, 1) * 8, 0);
    for (size_t b = 0; b < n_pos; b++) {
        int *rec = &r.block[b * 8];
        const int type = bp[b], ptr = bp[b + 1];
        rec[0] = type;
        rec[1] = ptr;
        if (type == 1) {
            // full cube: the six material words travel with the entry
            if (ptr >= 0 && (size_t)ptr + 5 < mp.size()) {
                for (int k = 0; k < 6; k++) rec[2 + k] = mp[(size_t)ptr + k];
            } else {
                rec[0] = 0x7FFFFFFF;
            }
        } else if (type == 2) {
            if (!model(aabbs, n_aabbs, ptr, 13, r.aabb, aabb_at, rec[1])) rec[0] = 0x7FFFFFFF;
        } else if (type == 3) {
            if (!model(quads, n_quads, ptr, 15, r.quad, quad_at, rec[1])) rec[0] = 0x7FFFFFFF;
        }
    }
    if (r.quad.empty()) r.quad.assign(4, 0);
    if (r.aabb.empty()) r.aabb.assign(4, 0);
    return r;
}

}  // namespace
