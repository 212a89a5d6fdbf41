// Native meto tokenizer, LR_ABSCO, LR and CLERS backends, ENCODE side (mesh -> token stream) behind the C ABI.
//
// Stands in for the pybind module `_meto` of the reference on the training-data side:
//   Mesh::Mesh                 meto/include/meto/mesh.h:172-278     (quantise, half-edges, twins, boundary marks, ordering heuristics)
//   Engine_LR_ABSCO::encode    meto/include/meto/engine_lr_absco.h:66-220 (EdgeBreaker-style traversal, L / R / BOM ops + absolute coords)
//   Engine_LR::encode          meto/include/meto/engine_lr.h:54-169 (same traversal; coordinates as residuals of the parallelogram
//                              prediction, out-of-range residual -> token -1; a split always goes right first; a pending sub-mesh is
//                              opened even if its gate face was reached another way in the meantime, so faces can repeat)
// Same token stream, face order and face types as the reference, bit for bit (tests/test_meto_cpu.py against goldens produced by the
// compiled reference).  Differences by design: index-based flat arrays instead of a pointer graph (no per-element new/delete), the
// traversal is ITERATIVE with an explicit stack of pending sub-meshes (the reference recurses once per face: stack depth = faces per
// sub-mesh), a sorted edge table instead of std::map for twin lookup, re-entrant (no engine-held state).
//
// Ordering rules that must be reproduced exactly because they decide the token stream:
//   * vertex quantiser  min(int((x + 1) * bins / 2), bins - 1) in float arithmetic (mesh.h:31-35);
//   * a face's three half-edges are std::sort'ed with "boundary edge first, then by distance between the two opposite vertices"
//     (mesh.h:116-121; the comparator is not a strict weak order for two boundary edges — we call std::sort with the same predicate
//     on the same initial order, which is what the reference does);
//   * faces are sorted by centre (y, z, x), connected components are labelled by BFS in that order, then faces are sorted again by
//     (component, centre) (mesh.h:235-277).
#include "../../include/edgerunner_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <utility>
#include <vector>

namespace {

enum : int32_t { kLeft = 0, kRight = 1, kBegin = 2, kNumOps = 3 };

struct QVert { int x, y, z; int mark; };
struct HalfEdge {
    int v, s, e;      // opposite vertex, start, end (vertex ids)
    int face;
    int next, prev;   // half-edge ids inside the face
    int twin;         // opposite half-edge id or -1
};
struct Face {
    int he[3];        // half-edge ids in heuristic order (he[0] starts a sub-mesh)
    int vid[3];
    float cx, cy, cz;
    int index;        // position in the input triangle list
    int comp;
    int mark;
};

inline int quantise(float x, int bins) { return std::min(int((x + 1) * bins / 2), bins - 1); }

struct Builder {
    std::vector<QVert> V;
    std::vector<HalfEdge> H;
    std::vector<Face> F;
    std::vector<int> order;     // face ids in traversal-start order (after the two sorts)

    float opp_dist(int h) const {   // |v(h) - v(twin(h))| on the integer grid, float arithmetic like the reference
        const QVert& a = V[H[h].v];
        const QVert& b = V[H[H[h].twin].v];
        const float dx = float(b.x - a.x), dy = float(b.y - a.y), dz = float(b.z - a.z);
        return std::sqrt(dx * dx + dy * dy + dz * dz);
    }
    bool he_less(int a, int b) const {
        if (H[a].twin < 0) return true;
        if (H[b].twin < 0) return false;
        return opp_dist(a) < opp_dist(b);
    }
    bool face_less(int a, int b) const {
        const Face &f = F[a], &g = F[b];
        if (f.comp != g.comp) return f.comp < g.comp;
        return f.cy < g.cy || (f.cy == g.cy && f.cz < g.cz) || (f.cy == g.cy && f.cz == g.cz && f.cx < g.cx);
    }

    void build(const float* verts, int64_t nv, const int32_t* tris, int64_t nf, int bins) {
        V.resize(nv);
        for (int64_t i = 0; i < nv; ++i)
            V[i] = QVert{quantise(verts[3 * i], bins), quantise(verts[3 * i + 1], bins), quantise(verts[3 * i + 2], bins), 0};
        H.resize(3 * nf);
        F.resize(nf);
        // edge table: (min, max, half-edge id) sorted by key, ties in creation order == the order std::map would have met them
        struct EdgeRef { int a, b, h; };
        std::vector<EdgeRef> edges(3 * nf);
        for (int64_t f = 0; f < nf; ++f) {
            Face& fc = F[f];
            fc.index = (int)f; fc.comp = -1; fc.mark = 0;
            for (int j = 0; j < 3; ++j) {
                const int h = int(3 * f + j);
                HalfEdge& e = H[h];
                e.v = tris[3 * f + j]; e.s = tris[3 * f + (j + 1) % 3]; e.e = tris[3 * f + (j + 2) % 3];
                e.face = (int)f; e.next = int(3 * f + (j + 1) % 3); e.prev = int(3 * f + (j + 2) % 3); e.twin = -1;
                fc.he[j] = h; fc.vid[j] = e.v;
                edges[h] = EdgeRef{std::min(e.s, e.e), std::max(e.s, e.e), h};
            }
            const QVert &a = V[fc.vid[0]], &b = V[fc.vid[1]], &c = V[fc.vid[2]];
            fc.cx = float(float(a.x + b.x + c.x) / 3.0); fc.cy = float(float(a.y + b.y + c.y) / 3.0); fc.cz = float(float(a.z + b.z + c.z) / 3.0);
        }
        std::stable_sort(edges.begin(), edges.end(), [](const EdgeRef& x, const EdgeRef& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
        for (size_t i = 0; i < edges.size();) {
            size_t j = i;
            while (j < edges.size() && edges[j].a == edges[i].a && edges[j].b == edges[i].b) ++j;
            // first two half-edges of an edge become twins; a third (non-manifold) one stays a border edge, like the reference
            if (j - i >= 2) { H[edges[i].h].twin = edges[i + 1].h; H[edges[i + 1].h].twin = edges[i].h; }
            i = j;
        }
        // boundary vertices start out "visited"; then order each face's half-edges
        for (int64_t f = 0; f < nf; ++f) {
            for (int j = 0; j < 3; ++j) {
                const HalfEdge& e = H[F[f].he[j]];
                if (e.twin < 0) { V[e.s].mark = 1; V[e.e].mark = 1; }
            }
            std::sort(F[f].he, F[f].he + 3, [this](int a, int b) { return he_less(a, b); });
        }
        order.resize(nf);
        for (int64_t f = 0; f < nf; ++f) order[f] = (int)f;
        std::sort(order.begin(), order.end(), [this](int a, int b) { return face_less(a, b); });
        int ncomp = 0;
        for (int64_t i = 0; i < nf; ++i) {
            if (F[order[i]].comp != -1) continue;
            ++ncomp;
            std::queue<int> q;
            q.push(order[i]);
            while (!q.empty()) {
                const int f = q.front(); q.pop();
                if (F[f].comp != -1) continue;
                F[f].comp = ncomp;
                for (int j = 0; j < 3; ++j) {
                    const int tw = H[F[f].he[j]].twin;
                    if (tw >= 0 && F[H[tw].face].comp == -1) q.push(H[tw].face);
                }
            }
        }
        std::sort(order.begin(), order.end(), [this](int a, int b) { return face_less(a, b); });
    }

    void flip(int f) {   // reverse the orientation of a face: swap start/end and next/prev of its three half-edges
        for (int j = 0; j < 3; ++j) {
            HalfEdge& e = H[F[f].he[j]];
            std::swap(e.s, e.e);
            std::swap(e.next, e.prev);
        }
    }
};

struct Emitter {
    std::vector<int32_t> tok, ord, typ;
    int bins; bool lr;
    void abs3(const QVert& v) {                     // a vertex written as such: +3 (LR_ABSCO) or + bins + 3 (LR)
        const int off = lr ? bins + kNumOps : kNumOps;
        tok.push_back(v.x + off); tok.push_back(v.y + off); tok.push_back(v.z + off);
    }
    int32_t rel(int d) const { return (d < -bins || d >= bins) ? -1 : d + bins + kNumOps; }   // LR residual / delta token
    void rel3(int dx, int dy, int dz) { tok.push_back(rel(dx)); tok.push_back(rel(dy)); tok.push_back(rel(dz)); }
};

// Engine_CLERS::encode (meto/include/meto/engine_clers.h:66-183): the classic EdgeBreaker walk.  Per face: (residual of the parallelogram
// prediction, 3 tokens — not for the first face of a sub-mesh) then the op: C (tip unseen: mark it, go right), E (both neighbours seen: the
// strip ends), L (left seen: go right), R (right seen: go left), S (neither: go right, then left).  The reference recurses per face; here an
// explicit LIFO of gates gives the same depth-first order (S pushes left, then right).  Like Engine_LR it re-enters a gate even when its face
// was reached another way in the meantime, so faces can repeat.
void clers_encode(Builder& m, int bins, int64_t n_faces, std::vector<int32_t>& tok, std::vector<int32_t>& ord, std::vector<int32_t>& typ) {
    enum : int32_t { C_ = 0, L_ = 1, E_ = 2, R_ = 3, S_ = 4, BOM_ = 5, EOM_ = 6, NUM_ = 7 };
    const int off = 2 * bins + NUM_;
    auto visited_face = [&](int h) { return m.F[m.H[h].face].mark != 0; };
    std::vector<int> gates;
    for (int64_t i = 0; i < n_faces; ++i) {
        if (m.F[m.order[i]].mark) continue;
        int c = m.F[m.order[i]].he[0];
        tok.push_back(BOM_);
        {
            const QVert &qv = m.V[m.H[c].v], &qs = m.V[m.H[c].s], &qe = m.V[m.H[c].e];
            tok.push_back(qv.x + off); tok.push_back(qv.y + off); tok.push_back(qv.z + off);
            tok.push_back(qs.x - qv.x + off); tok.push_back(qs.y - qv.y + off); tok.push_back(qs.z - qv.z + off);
            tok.push_back(qe.x - qs.x + off); tok.push_back(qe.y - qs.y + off); tok.push_back(qe.z - qs.z + off);
        }
        m.V[m.H[c].s].mark = 1; m.V[m.H[c].e].mark = 1;
        gates.clear();
        gates.push_back(c);
        bool init = true;
        while (!gates.empty()) {
            c = gates.back();
            gates.pop_back();
            if (c < 0) { init = false; continue; }          // no face across that edge (the reference would dereference NULL)
            m.F[m.H[c].face].mark = 1;
            ord.push_back(m.F[m.H[c].face].index);
            if (!init) {
                const HalfEdge& tw = m.H[m.H[c].twin];
                if (!(m.H[c].s == tw.e && m.H[c].e == tw.s)) m.flip(m.H[c].face);
                const HalfEdge& hf = m.H[c];
                const QVert &qv = m.V[hf.v], &qo = m.V[m.H[hf.twin].v], &qn = m.V[m.H[hf.next].v], &qp = m.V[m.H[hf.prev].v];
                tok.push_back(qv.x + qo.x - qn.x - qp.x + off); tok.push_back(qv.y + qo.y - qn.y - qp.y + off); tok.push_back(qv.z + qo.z - qn.z - qp.z + off);
            }
            init = false;
            const HalfEdge& h = m.H[c];
            const bool tip_seen = m.V[h.v].mark != 0;
            const int left_gate = m.H[h.prev].twin, right_gate = m.H[h.next].twin;
            const bool left_seen = left_gate < 0 || visited_face(left_gate);
            const bool right_seen = right_gate < 0 || visited_face(right_gate);
            if (!tip_seen) { tok.push_back(C_); typ.push_back(C_); m.V[h.v].mark = 1; gates.push_back(right_gate); }
            else if (left_seen && right_seen) { tok.push_back(E_); typ.push_back(E_); }
            else if (left_seen) { tok.push_back(L_); typ.push_back(L_); gates.push_back(right_gate); }
            else if (right_seen) { tok.push_back(R_); typ.push_back(R_); gates.push_back(left_gate); }
            else { tok.push_back(S_); typ.push_back(S_); gates.push_back(left_gate); gates.push_back(right_gate); }
        }
        tok.push_back(EOM_);
    }
}

}  // namespace

extern "C" int er_meto_encode(int32_t backend, int32_t discrete_bins, const float* verts, int64_t n_verts, const int32_t* faces, int64_t n_faces,
                              int32_t* tokens, int64_t tokens_cap, int32_t* face_order, int32_t* face_type, int64_t faces_cap,
                              int64_t* n_tokens, int64_t* n_faces_out) {
    if ((backend != ER_METO_LR_ABSCO && backend != ER_METO_LR && backend != ER_METO_CLERS) || discrete_bins <= 0 || n_verts < 0 || n_faces < 0 ||
        (n_faces > 0 && (!verts || !faces)) || !tokens || !face_order || !face_type || !n_tokens || !n_faces_out)
        return ER_ERR_INVALID;
    for (int64_t i = 0; i < 3 * n_faces; ++i)
        if (faces[i] < 0 || faces[i] >= n_verts) return ER_ERR_INVALID;   // the reference would read out of bounds
    const bool lr = backend == ER_METO_LR;
    Builder m;
    m.build(verts, n_verts, faces, n_faces, discrete_bins);
    Emitter out;
    out.bins = discrete_bins; out.lr = lr;
    out.tok.reserve(10 * n_faces + 16); out.ord.reserve(n_faces + 16); out.typ.reserve(n_faces + 16);
    if (backend == ER_METO_CLERS) {
        clers_encode(m, discrete_bins, n_faces, out.tok, out.ord, out.typ);
        *n_tokens = (int64_t)out.tok.size();
        *n_faces_out = (int64_t)out.ord.size();
        if ((int64_t)out.tok.size() > tokens_cap || (int64_t)out.ord.size() > faces_cap) return ER_ERR_CAPACITY;
        std::copy(out.tok.begin(), out.tok.end(), tokens);
        std::copy(out.ord.begin(), out.ord.end(), face_order);
        std::copy(out.typ.begin(), out.typ.end(), face_type);
        return ER_OK;
    }
    std::vector<int> pending;   // sub-meshes still to be opened (LIFO == the reference's recursion order)
    auto visited_face = [&](int h) { return m.F[m.H[h].face].mark != 0; };
    for (int64_t i = 0; i < n_faces; ++i) {
        if (m.F[m.order[i]].mark) continue;
        pending.push_back(m.F[m.order[i]].he[0]);
        while (!pending.empty()) {
            int c = pending.back();
            pending.pop_back();
            if (!lr && visited_face(c)) continue;             // hole / handle: the face was reached another way (LR opens it regardless)
            // ---- open a sub-mesh at gate c: BOM + three vertices (absolute; LR: first absolute, then two deltas) ----
            out.tok.push_back(kBegin);
            {
                const QVert &qv = m.V[m.H[c].v], &qs = m.V[m.H[c].s], &qe = m.V[m.H[c].e];
                out.abs3(qv);
                if (lr) { out.rel3(qs.x - qv.x, qs.y - qv.y, qs.z - qv.z); out.rel3(qe.x - qs.x, qe.y - qs.y, qe.z - qs.z); }
                else { out.abs3(qs); out.abs3(qe); }
            }
            m.V[m.H[c].s].mark = 1; m.V[m.H[c].e].mark = 1;
            bool first = true;
            for (;;) {                                        // one iteration per face (the reference recurses here)
                HalfEdge& hc = m.H[c];
                m.F[hc.face].mark = 1;
                out.ord.push_back(m.F[hc.face].index);
                if (!first) {
                    const HalfEdge& tw = m.H[hc.twin];       // the gate we came through
                    if (!(hc.s == tw.e && hc.e == tw.s)) m.flip(hc.face);   // inconsistent orientation: repair it
                    const HalfEdge& hf = m.H[c];             // (after the flip)
                    const QVert& qv = m.V[hf.v];
                    if (lr) {                                // residual of the parallelogram prediction  twin.v' = next.v + prev.v - twin.v
                        const QVert &qo = m.V[m.H[hf.twin].v], &qn = m.V[m.H[hf.next].v], &qp = m.V[m.H[hf.prev].v];
                        out.rel3(qv.x + qo.x - qn.x - qp.x, qv.y + qo.y - qn.y - qp.y, qv.z + qo.z - qn.z - qp.z);
                    } else {
                        out.abs3(qv);
                    }
                }
                first = false;
                const HalfEdge& h = m.H[c];                   // (re-read: flip may have changed next / prev)
                const bool tip_seen = m.V[h.v].mark != 0;
                const int left_gate = m.H[h.prev].twin, right_gate = m.H[h.next].twin;
                const bool left_seen = left_gate < 0 || visited_face(left_gate);
                const bool right_seen = right_gate < 0 || visited_face(right_gate);
                if (!tip_seen) {                              // "C": new vertex, continue to the right
                    m.V[h.v].mark = 1;
                    out.tok.push_back(kLeft); out.typ.push_back(kLeft);
                    c = right_gate;
                } else if (left_seen && right_seen) {         // "E": this strip ends
                    out.typ.push_back(kBegin);
                    break;
                } else if (left_seen) {
                    out.tok.push_back(kLeft); out.typ.push_back(kLeft);
                    c = right_gate;
                } else if (right_seen) {
                    out.tok.push_back(kRight); out.typ.push_back(kRight);
                    c = left_gate;
                } else if (lr) {                              // "S" (LR): always to the right; the left side becomes a new sub-mesh later
                    out.tok.push_back(kLeft); out.typ.push_back(kLeft);
                    pending.push_back(left_gate);
                    c = right_gate;
                } else {                                      // "S": split — walk the shorter boundary loop first
                    int len_left = 0, len_right = 0;
                    for (int cur = right_gate;;) {
                        ++len_left;
                        cur = m.H[cur].next;
                        while (m.H[cur].twin >= 0 && !visited_face(m.H[cur].twin)) cur = m.H[m.H[cur].twin].next;
                        if (cur == right_gate) break;
                    }
                    for (int cur = left_gate;;) {
                        ++len_right;
                        cur = m.H[cur].prev;
                        while (m.H[cur].twin >= 0 && !visited_face(m.H[cur].twin)) cur = m.H[m.H[cur].twin].prev;
                        if (cur == left_gate) break;
                    }
                    if (len_left < len_right) {
                        out.tok.push_back(kLeft); out.typ.push_back(kLeft);
                        pending.push_back(left_gate);
                        c = right_gate;
                    } else {
                        out.tok.push_back(kRight); out.typ.push_back(kRight);
                        pending.push_back(right_gate);
                        c = left_gate;
                    }
                }
            }
        }
    }
    *n_tokens = (int64_t)out.tok.size();
    *n_faces_out = (int64_t)out.ord.size();
    if ((int64_t)out.tok.size() > tokens_cap || (int64_t)out.ord.size() > faces_cap) return ER_ERR_CAPACITY;   // counts are valid: retry
    std::copy(out.tok.begin(), out.tok.end(), tokens);
    std::copy(out.ord.begin(), out.ord.end(), face_order);
    std::copy(out.typ.begin(), out.typ.end(), face_type);
    return ER_OK;
}
