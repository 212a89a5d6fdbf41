// Native mesh post-processing behind the C ABI: the clean-up tail of save_mesh.
//
// The reference delegates it to third-party trimesh (core/provider.py:55-58: merge_vertices, update_faces(unique_faces()),
// fix_normals) — un-vendored, so parity is on the documented behaviour, not on trimesh's bits:
//   1. merge vertices whose coordinates agree after rounding to `digits` decimals (trimesh: tol.merge = 1e-8 -> 8 digits);
//      vertices keep the order of their first occurrence, faces are re-indexed;
//   2. drop faces whose vertex SET already occurred (first occurrence wins);
//   3. make the winding consistent inside each edge-connected component (breadth-first over shared edges), then invert every
//      component whose signed volume is negative (outward normals for closed components).
// Flat arrays in / out, no allocation visible to the caller, O((V + F) log) time.
#include "../../include/edgerunner_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

extern "C" int er_mesh_clean(const double* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces, int32_t digits,
                             double* verts_out, int32_t* faces_out, int64_t* n_verts_out, int64_t* n_faces_out) {
    if (n_verts < 0 || n_faces < 0 || (n_verts > 0 && !verts) || (n_faces > 0 && !faces) || !verts_out || !faces_out || !n_verts_out || !n_faces_out ||
        digits < 0 || digits > 15)
        return ER_ERR_INVALID;
    for (int64_t i = 0; i < 3 * n_faces; ++i)
        if (faces[i] < 0 || faces[i] >= n_verts) return ER_ERR_INVALID;

    // ---- 1. merge vertices on the rounded grid ----
    const double scale = std::pow(10.0, digits);
    std::map<std::array<long long, 3>, int32_t> seen;
    std::vector<int32_t> remap(n_verts);
    int64_t nv = 0;
    for (int64_t i = 0; i < n_verts; ++i) {
        const std::array<long long, 3> key = {std::llround(verts[3 * i] * scale), std::llround(verts[3 * i + 1] * scale), std::llround(verts[3 * i + 2] * scale)};
        auto it = seen.find(key);
        if (it == seen.end()) {
            it = seen.emplace(key, (int32_t)nv).first;
            for (int k = 0; k < 3; ++k) verts_out[3 * nv + k] = verts[3 * i + k];
            ++nv;
        }
        remap[i] = it->second;
    }

    // ---- 2. unique faces (by vertex set), first occurrence wins ----
    std::map<std::array<int32_t, 3>, int> face_seen;
    std::vector<std::array<int32_t, 3>> F;
    F.reserve(n_faces);
    for (int64_t i = 0; i < n_faces; ++i) {
        std::array<int32_t, 3> f = {remap[faces[3 * i]], remap[faces[3 * i + 1]], remap[faces[3 * i + 2]]};
        std::array<int32_t, 3> key = f;
        std::sort(key.begin(), key.end());
        if (face_seen.emplace(key, 1).second) F.push_back(f);
    }
    const int64_t nf = (int64_t)F.size();

    // ---- 3. consistent winding per component, then outward orientation ----
    struct EdgeUse { int32_t a, b, face; };                    // undirected key (a < b)
    std::vector<EdgeUse> uses;
    uses.reserve(3 * nf);
    for (int64_t i = 0; i < nf; ++i)
        for (int k = 0; k < 3; ++k) {
            const int32_t u = F[i][k], v = F[i][(k + 1) % 3];
            uses.push_back({std::min(u, v), std::max(u, v), (int32_t)i});
        }
    std::stable_sort(uses.begin(), uses.end(), [](const EdgeUse& x, const EdgeUse& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
    std::vector<int64_t> first_use(3 * nf, -1), last_use(3 * nf, -1);   // per (face, corner): range of uses of that edge
    for (size_t i = 0; i < uses.size();) {
        size_t j = i;
        while (j < uses.size() && uses[j].a == uses[i].a && uses[j].b == uses[i].b) ++j;
        for (size_t q = i; q < j; ++q) {
            const int32_t f = uses[q].face;
            for (int k = 0; k < 3; ++k) {
                const int32_t u = F[f][k], v = F[f][(k + 1) % 3];
                if (std::min(u, v) == uses[i].a && std::max(u, v) == uses[i].b) { first_use[3 * f + k] = (int64_t)i; last_use[3 * f + k] = (int64_t)j; }
            }
        }
        i = j;
    }
    auto has_directed = [&](int32_t f, int32_t u, int32_t v) {
        for (int k = 0; k < 3; ++k) if (F[f][k] == u && F[f][(k + 1) % 3] == v) return true;
        return false;
    };
    std::vector<int32_t> comp(nf, -1);
    std::vector<int32_t> stack;
    int32_t ncomp = 0;
    for (int64_t s = 0; s < nf; ++s) {
        if (comp[s] >= 0) continue;
        comp[s] = ncomp;
        stack.push_back((int32_t)s);
        while (!stack.empty()) {
            const int32_t i = stack.back();
            stack.pop_back();
            for (int k = 0; k < 3; ++k) {
                const int32_t u = F[i][k], v = F[i][(k + 1) % 3];
                for (int64_t q = first_use[3 * i + k]; q < last_use[3 * i + k]; ++q) {
                    const int32_t j = uses[q].face;
                    if (j == i || comp[j] >= 0) continue;
                    if (has_directed(j, u, v)) {                              // same direction on the shared edge: the neighbour is flipped
                        std::swap(F[j][0], F[j][2]);                          // (a,b,c) -> (c,b,a): corner 0 now starts edge (c,b), corner 1 edge (b,a)
                        std::swap(first_use[3 * j], first_use[3 * j + 1]);
                        std::swap(last_use[3 * j], last_use[3 * j + 1]);
                    }
                    comp[j] = ncomp;
                    stack.push_back(j);
                }
            }
        }
        ++ncomp;
    }
    std::vector<double> vol6(ncomp, 0.0);
    for (int64_t i = 0; i < nf; ++i) {
        const double* a = verts_out + 3 * F[i][0]; const double* b = verts_out + 3 * F[i][1]; const double* c = verts_out + 3 * F[i][2];
        vol6[comp[i]] += a[0] * (b[1] * c[2] - b[2] * c[1]) + a[1] * (b[2] * c[0] - b[0] * c[2]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
    }
    for (int64_t i = 0; i < nf; ++i) {
        if (vol6[comp[i]] < 0) std::swap(F[i][0], F[i][2]);
        faces_out[3 * i] = F[i][0]; faces_out[3 * i + 1] = F[i][1]; faces_out[3 * i + 2] = F[i][2];
    }
    *n_verts_out = nv;
    *n_faces_out = nf;
    return ER_OK;
}
