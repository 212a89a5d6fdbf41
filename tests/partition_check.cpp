// Native check of the decode kernel's work partition (edgerunner_b200/csrc/decode_partition.h), compiled with g++ by
// tests/test_abi_cpu.py::test_decode_partition_invariants.  Exits 0 and prints "ok <cases>" when every invariant holds.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../edgerunner_b200/csrc/decode_partition.h"

#define REQUIRE(c) do { if (!(c)) { std::printf("FAILED %s (line %d): R=%d G=%u b=%u H=%d S=%d L=%d hc=%d\n", #c, __LINE__, R, G, b, H, S, L, hc); return 1; } } while (0)

int main() {
    long cases = 0;
    int R = 0, H = 0, S = 0, L = 0, hc = 0;
    unsigned G = 0, b = 0;
    // ---- row ranges: tile [0, R) in CTA order, on fp16-pair boundaries when R is even, balanced to within one pair ----
    const int Rs[] = {4608, 1536, 6144, 518, 576, 192, 768, 134, 2304, 768 * 3, 3072, 1030, 7, 1};
    const unsigned Gs[] = {148, 132, 108, 64, 16, 1};
    for (int Ri : Rs) for (unsigned Gi : Gs) {
        R = Ri; G = Gi;
        int next = 0, lo = 1 << 30, hi = 0;
        for (b = 0; b < G; ++b) {
            const er::RowRange rr = er::cta_rows_of(R, b, G);
            REQUIRE(rr.r0 == next && rr.r1 >= rr.r0);
            if (!(R & 1)) REQUIRE(rr.r0 % 2 == 0 && rr.r1 % 2 == 0);
            next = rr.r1;
            lo = rr.r1 - rr.r0 < lo ? rr.r1 - rr.r0 : lo;
            hi = rr.r1 - rr.r0 > hi ? rr.r1 - rr.r0 : hi;
            ++cases;
        }
        REQUIRE(next == R);
        REQUIRE(hi - lo <= ((R & 1) ? 1 : 2));
        REQUIRE(hi <= (R + (int)G - 1) / (int)G + 1);          // the bound engine.cu uses for max_units
    }
    // ---- attention splits: per head, blocks and keys tiled exactly once; one owner of the new key; scratch large enough ----
    const int HS[][2] = {{16, 9}, {2, 16}, {12, 12}, {8, 16}, {16, 1}, {1, 16}};
    for (const auto& hs : HS) for (hc = 0; hc <= 7; ++hc) for (L = 1; L <= 20000; L += (L < 300 ? 1 : 37)) {
        H = hs[0]; S = hs[1]; G = (unsigned)(H * S + 4);
        const int nblk = (L + 31) >> 5, nkb = (20000 + 64 + 31) / 32;
        const int sc_len = er::score_scratch_len(nkb, S, 518);
        for (int h = 0; h < H; ++h) {
            int next_block = 0, next_key = 0, owners = 0;
            for (int s = 0; s < S; ++s) {
                b = (unsigned)(h * S + s);
                er::AttnRange a;
                REQUIRE(er::attn_range_of(H, S, hc, L, b, a));
                REQUIRE(a.h == h && a.b0 == next_block && a.b1 >= a.b0 && a.b1 <= nblk);
                REQUIRE(a.k0 == a.b0 * 32 && a.k1 >= a.k0);
                if (a.b1 > a.b0) { REQUIRE(a.k0 == next_key && a.k1 <= L && a.k1 > a.k0); next_key = a.k1; }
                else REQUIRE(a.k1 == a.k0);                                  // empty split: no keys, no blocks
                next_block = a.b1;
                owners += a.is_new;
                REQUIRE((a.b1 - a.b0) * 32 + 33 <= sc_len);                  // scores of the old blocks + the new key's slot
                REQUIRE(a.b1 == a.b0 || (a.k1 - a.k0) > (a.b1 - a.b0 - 1) * 32);   // only the last block of a split may be partial
                ++cases;
            }
            REQUIRE(next_block == nblk && next_key == L && owners == 1);
        }
        er::AttnRange a;
        b = (unsigned)(H * S);
        REQUIRE(!er::attn_range_of(H, S, hc, L, b, a));                       // CTAs beyond H * S take no part in attention
    }
    std::printf("ok %ld\n", cases);
    return 0;
}
