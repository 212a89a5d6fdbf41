// Work partition of the persistent decode kernel: which output rows and which (head, KV split) a CTA owns.
// Plain integer functions shared by the producer and the consumer side of decode_kernel.cu and by the host (engine.cu sizes the
// score scratch from them; tests/partition_check.cpp checks their invariants natively: row ranges tile [0, R) on pair boundaries,
// the splits of a head tile its keys and blocks exactly once, exactly one split owns the new key).
#pragma once

#if defined(__CUDACC__)
#define ER_HD __host__ __device__ __forceinline__
#else
#define ER_HD inline
#endif

namespace er {

#if defined(__CUDA_ARCH__)
ER_HD int er_imin(int a, int b) { return min(a, b); }
ER_HD int er_imax(int a, int b) { return max(a, b); }
#else
ER_HD int er_imin(int a, int b) { return a < b ? a : b; }
ER_HD int er_imax(int a, int b) { return a > b ? a : b; }
#endif

struct RowRange { int r0, r1; };
// rows [r0, r1) of an R-row phase for CTA b of g.  R * g < 2^31 (checked by the host): 32-bit arithmetic, no division call.
ER_HD RowRange cta_rows_of(int R, unsigned b, unsigned g) {
    RowRange rr;
#ifdef ER_ODD_ROWS
    if (true) {
#else
    if (R & 1) {
#endif
        rr.r0 = (int)(((unsigned)R * b) / g);
        rr.r1 = (int)(((unsigned)R * (b + 1)) / g);
    } else {   // whole fp16 PAIRS of rows per CTA: one 8-byte exchange word never has two writers
        rr.r0 = 2 * (int)(((unsigned)(R >> 1) * b) / g);
        rr.r1 = 2 * (int)(((unsigned)(R >> 1) * (b + 1)) / g);
    }
    return rr;
}

struct AttnRange { int h, b0, b1, k0, k1, is_new; };   // old keys [k0,k1) in K blocks [b0,b1); is_new: this CTA also owns key L
// CTA b = (head b / S, split b % S) for b < H * S; L = keys already cached (the new key has index L)
ER_HD bool attn_range_of(int H, int S, int split_handicap, int L, unsigned b, AttnRange& a) {
    if ((int)b >= H * S) return false;
    a.h = b / S;                                     // (unsigned arithmetic, like blockIdx.x / S)
    const int s = b % S;
    const int nblk = (L + 31) >> 5;                  // blocks holding old keys 0..L-1
    // the last split also owns the new key and (being the last to finish) usually merges the head: that fixed work is worth
    // about split_handicap (<= 7, see engine.cu sc_len) blocks of streaming, so it gets that many fewer blocks
    const int bps = (nblk + split_handicap + S - 1) / S;
    a.b0 = er_imin(s * bps, nblk);
    a.b1 = er_imin(a.b0 + bps, nblk);
    a.k0 = a.b0 * 32;
    a.k1 = er_imax(a.k0, er_imin(a.b1 * 32, L));    // empty splits: b0 == b1 == nblk, k0 may exceed L
    a.is_new = (s == S - 1);
    return true;
}

// floats of score scratch a CTA needs: the keys of its widest split, the new key's slot, and the sampler's V scores
ER_HD int score_scratch_len(int nkb, int S, int V) {
    const int keys = ((nkb + 7 + S - 1) / S) * 32 + 64;     // 7 = largest split handicap er_decode accepts
    const int v4 = (V + 3) / 4 * 4;
    return keys > v4 ? keys : v4;
}

}  // namespace er
