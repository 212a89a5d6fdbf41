// Native meto detokenizer (LR_ABSCO, LR and CLERS backends) behind the C ABI — the tail of LMM.generate.
//
// Stands in for the pybind module `_meto` of the reference (meto/src/bindings.cpp:18-28) on the decode path:
//   Engine_LR_ABSCO::decode   meto/include/meto/engine_lr_absco.h:223-295   (absolute coordinates, the ArAE / DiT presets)
//   Engine_LR::decode         meto/include/meto/engine_lr.h:171-254         (parallelogram-predicted residuals, Options.meto_backend = 'LR')
//   Vertex::undiscrete        meto/include/meto/mesh.h:39-45
// Differences by design: flat caller-owned arrays instead of vector<vector<>> + Python lists (the reference's
// conversion dominates its run time), no engine-held state (re-entrant, thread-safe), no exceptions.
// Semantics kept bit-for-bit, including the quirks: a stream is cut at the first structural error (truncated
// triangle / vertex, coordinate where an operator is expected); coordinates are not range-checked; every face
// gets fresh vertices (no sharing); face_type always ends with a BOM entry.
#include "../../include/edgerunner_b200.h"

namespace {

enum : int32_t { kLeft = 0, kRight = 1, kBegin = 2, kNumOps = 3 };

struct Corner { int32_t q[3]; int32_t id; };

inline float dequant(int32_t q, int32_t bins) {
    // bin centre in [-1, 1]; evaluated in double like the reference, then narrowed to float
    return static_cast<float>((static_cast<double>(static_cast<float>(q)) + 0.5) / bins * 2 - 1);
}

struct Sink {
    float* v; int32_t* f; int32_t* t; int32_t bins; bool lr;
    int64_t nv = 0, nf = 0, nt = 0;
    // token -> integer: LR_ABSCO stores the coordinate itself (+3); LR stores a signed residual shifted by bins + 3, and passes
    // negative tokens (its out-of-range marker -1) through unchanged (engine_lr.h:60-63)
    int32_t value(int32_t tok) const { return lr ? (tok < 0 ? tok : tok - bins - kNumOps) : tok - kNumOps; }
    // c = base + residual(tok) (LR) or c = coordinate(tok) (LR_ABSCO); `base` is the parallelogram prediction
    void vertex(Corner& c, const int32_t* tok, const int32_t* base = nullptr) {
        for (int k = 0; k < 3; ++k) {
            c.q[k] = value(tok[k]) + (base ? base[k] : 0);
            v[3 * nv + k] = dequant(c.q[k], bins);
        }
        c.id = static_cast<int32_t>(nv++);
    }
    void face(const Corner& a, const Corner& b, const Corner& c) {
        f[3 * nf] = a.id; f[3 * nf + 1] = b.id; f[3 * nf + 2] = c.id; ++nf;
    }
};

// Engine_CLERS::decode (meto/include/meto/engine_clers.h:186-283): ops C 0, L 1, E 2, R 3, S 4, BOM 5, EOM 6; coordinates are integers
// offset by 2 * bins + 7: absolute for the first corner of a sub-mesh, deltas for its second / third corner, parallelogram residuals after
// that.  C / L / S put the new triangle across the (tip, right) edge, R across (left, tip); S remembers the gate, and an E that is not
// followed by EOM pops it and continues to the left.  Quirks kept: the stream is cut at a truncated triangle / vertex or at a coordinate
// where an operator is expected; face_type gets E for every BOM but the first, E for a popped E, nothing for an E that ends a sub-mesh, and
// a final E.  Where the reference reads out of bounds (E as the very last token, E with an empty stack) this implementation stops.
int clers_decode(int32_t bins, const int32_t* tok, int64_t n, float* verts, int32_t* faces, int32_t* ftype, int64_t* n_verts, int64_t* n_faces,
                 int64_t* n_types) {
    enum : int32_t { C_ = 0, L_ = 1, E_ = 2, R_ = 3, S_ = 4, BOM_ = 5, EOM_ = 6, NUM_ = 7 };
    struct P { int32_t x, y, z, id; };
    const int32_t off = 2 * bins + NUM_;
    int64_t nv = 0, nf = 0, nt = 0;
    auto emit_v = [&](P& p) { p.id = (int32_t)nv; verts[3 * nv] = dequant(p.x, bins); verts[3 * nv + 1] = dequant(p.y, bins); verts[3 * nv + 2] = dequant(p.z, bins); ++nv; };
    auto emit_f = [&](const P& a, const P& b, const P& c) { faces[3 * nf] = a.id; faces[3 * nf + 1] = b.id; faces[3 * nf + 2] = c.id; ++nf; };
    P v0{}, v1{}, v2{};
    struct Gate { P a, b, c; };
    Gate* stack = nullptr; int64_t sp = 0, scap = 0;
    auto push = [&](const Gate& g) {
        if (sp == scap) { scap = scap ? 2 * scap : 64; Gate* ns = new Gate[scap]; for (int64_t k = 0; k < sp; ++k) ns[k] = stack[k]; delete[] stack; stack = ns; }
        stack[sp++] = g;
    };
    for (int64_t i = 0; i < n; ++i) {
        int32_t t = tok[i];
        if (t == BOM_) {
            if (i + 9 >= n) break;
            v0 = P{tok[i + 1] - off, tok[i + 2] - off, tok[i + 3] - off, 0}; emit_v(v0);
            v1 = P{v0.x + tok[i + 4] - off, v0.y + tok[i + 5] - off, v0.z + tok[i + 6] - off, 0}; emit_v(v1);
            v2 = P{v1.x + tok[i + 7] - off, v1.y + tok[i + 8] - off, v1.z + tok[i + 9] - off, 0}; emit_v(v2);
            emit_f(v0, v1, v2);
            if (i != 0) ftype[nt++] = E_;
            i += 9;
            continue;
        }
        if (t == EOM_) continue;
        if (t >= NUM_ || t < 0) break;                    // a coordinate where an operator is expected
        bool popped = false;
        if (t == E_) {
            if (i + 1 < n && tok[i + 1] == EOM_) continue;   // ends the sub-mesh
            if (i + 1 >= n || sp == 0) break;             // (the reference reads out of bounds here)
            ftype[nt++] = E_;
            popped = true;
            t = R_;
            const Gate g = stack[--sp];
            v0 = g.a; v1 = g.b; v2 = g.c;
        }
        if (i + 3 >= n) break;                            // truncated vertex
        const int32_t dx = tok[i + 1] - off, dy = tok[i + 2] - off, dz = tok[i + 3] - off;
        if (t == C_ || t == L_ || t == S_) {              // to the right: prediction v0 + v2 - v1
            P v{v0.x + v2.x - v1.x + dx, v0.y + v2.y - v1.y + dy, v0.z + v2.z - v1.z + dz, 0};
            emit_v(v);
            emit_f(v, v0, v2);
            if (t == S_) push(Gate{v0, v1, v2});
            v1 = v0; v0 = v;
        } else {                                          // R: to the left, prediction v0 + v1 - v2
            P v{v0.x + v1.x - v2.x + dx, v0.y + v1.y - v2.y + dy, v0.z + v1.z - v2.z + dz, 0};
            emit_v(v);
            emit_f(v, v1, v0);
            v2 = v0; v0 = v;
        }
        if (!popped) ftype[nt++] = t;
        i += 3;
    }
    ftype[nt++] = E_;
    delete[] stack;
    *n_verts = nv; *n_faces = nf; *n_types = nt;
    return ER_OK;
}

}  // namespace

extern "C" int er_meto_decode(int32_t backend, int32_t discrete_bins, const int32_t* tokens, int64_t n, float* verts, int32_t* faces,
                              int32_t* face_type, int64_t* n_verts, int64_t* n_faces, int64_t* n_types) {
    if ((backend != ER_METO_LR_ABSCO && backend != ER_METO_LR && backend != ER_METO_CLERS) || discrete_bins <= 0 || n < 0 || (n > 0 && !tokens) ||
        !verts || !faces || !face_type || !n_verts || !n_faces || !n_types)
        return ER_ERR_INVALID;
    if (backend == ER_METO_CLERS) return clers_decode(discrete_bins, tokens, n, verts, faces, face_type, n_verts, n_faces, n_types);
    const bool lr = backend == ER_METO_LR;
    Sink out{verts, faces, face_type, discrete_bins, lr};
    Corner tip{}, left{}, right{};   // the active gate: `tip` is the last emitted vertex
    int64_t i = 0;
    while (i < n) {
        const int32_t op = tokens[i];
        if (op == kBegin) {
            if (i + 9 >= n) break;                       // the reference requires one token beyond the 9 coordinates
            out.vertex(tip, tokens + i + 1);
            out.vertex(left, tokens + i + 4, lr ? tip.q : nullptr);       // LR: second / third corner as deltas from the previous one
            out.vertex(right, tokens + i + 7, lr ? left.q : nullptr);
            out.face(tip, left, right);
            if (i != 0) out.t[out.nt++] = kBegin;
            i += 10;
            continue;
        }
        if (op >= kNumOps) break;                        // coordinate where an operator is expected
        if (i + 3 >= n) break;                           // truncated vertex
        if (op == kLeft) {                               // new triangle across the (tip, right) edge; prediction tip + right - left
            Corner nv;
            const int32_t pred[3] = {tip.q[0] + right.q[0] - left.q[0], tip.q[1] + right.q[1] - left.q[1], tip.q[2] + right.q[2] - left.q[2]};
            out.vertex(nv, tokens + i + 1, lr ? pred : nullptr);
            out.face(nv, tip, right);
            left = tip; tip = nv;
        } else if (op == kRight) {                       // new triangle across the (left, tip) edge; prediction tip + left - right
            Corner nv;
            const int32_t pred[3] = {tip.q[0] + left.q[0] - right.q[0], tip.q[1] + left.q[1] - right.q[1], tip.q[2] + left.q[2] - right.q[2]};
            out.vertex(nv, tokens + i + 1, lr ? pred : nullptr);
            out.face(nv, left, tip);
            right = tip; tip = nv;
        }                                                // (negative ids: the reference emits nothing but records the type)
        out.t[out.nt++] = op;
        i += 4;
    }
    out.t[out.nt++] = kBegin;
    *n_verts = out.nv; *n_faces = out.nf; *n_types = out.nt;
    return ER_OK;
}
