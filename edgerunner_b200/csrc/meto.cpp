// Native meto detokenizer (LR_ABSCO and LR backends) behind the C ABI — the tail of LMM.generate.
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

}  // namespace

extern "C" int er_meto_decode(int32_t backend, int32_t discrete_bins, const int32_t* tokens, int64_t n, float* verts, int32_t* faces,
                              int32_t* face_type, int64_t* n_verts, int64_t* n_faces, int64_t* n_types) {
    if ((backend != ER_METO_LR_ABSCO && backend != ER_METO_LR) || discrete_bins <= 0 || n < 0 || (n > 0 && !tokens) || !verts || !faces ||
        !face_type || !n_verts || !n_faces || !n_types)
        return ER_ERR_INVALID;
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
