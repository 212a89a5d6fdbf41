// Native meto detokenizer (LR_ABSCO backend) behind the C ABI — the tail of LMM.generate.
//
// Stands in for the pybind module `_meto` of the reference (meto/src/bindings.cpp:25-28) on the decode path:
//   Engine_LR_ABSCO::decode   meto/include/meto/engine_lr_absco.h:223-295
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
    float* v; int32_t* f; int32_t* t; int32_t bins;
    int64_t nv = 0, nf = 0, nt = 0;
    void vertex(Corner& c, const int32_t* tok) {
        for (int k = 0; k < 3; ++k) {
            c.q[k] = tok[k] - kNumOps;
            v[3 * nv + k] = dequant(c.q[k], bins);
        }
        c.id = static_cast<int32_t>(nv++);
    }
    void face(const Corner& a, const Corner& b, const Corner& c) {
        f[3 * nf] = a.id; f[3 * nf + 1] = b.id; f[3 * nf + 2] = c.id; ++nf;
    }
};

}  // namespace

extern "C" int er_meto_decode(int32_t discrete_bins, const int32_t* tokens, int64_t n, float* verts, int32_t* faces, int32_t* face_type,
                              int64_t* n_verts, int64_t* n_faces, int64_t* n_types) {
    if (discrete_bins <= 0 || n < 0 || (n > 0 && !tokens) || !verts || !faces || !face_type || !n_verts || !n_faces || !n_types) return ER_ERR_INVALID;
    Sink out{verts, faces, face_type, discrete_bins};
    Corner tip{}, left{}, right{};   // the active gate: `tip` is the last emitted vertex
    int64_t i = 0;
    while (i < n) {
        const int32_t op = tokens[i];
        if (op == kBegin) {
            if (i + 9 >= n) break;                       // the reference requires one token beyond the 9 coordinates
            out.vertex(tip, tokens + i + 1);
            out.vertex(left, tokens + i + 4);
            out.vertex(right, tokens + i + 7);
            out.face(tip, left, right);
            if (i != 0) out.t[out.nt++] = kBegin;
            i += 10;
            continue;
        }
        if (op >= kNumOps) break;                        // coordinate where an operator is expected
        if (i + 3 >= n) break;                           // truncated vertex
        if (op == kLeft) {                               // new triangle across the (tip, right) edge
            Corner nv;
            out.vertex(nv, tokens + i + 1);
            out.face(nv, tip, right);
            left = tip; tip = nv;
        } else if (op == kRight) {                       // new triangle across the (left, tip) edge
            Corner nv;
            out.vertex(nv, tokens + i + 1);
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
