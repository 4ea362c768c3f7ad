// Host-only check of the compile-time cascade plans of the warp-streaming fused front (fused_front2.cuh):
// prints one line per plan and exits non-zero when an invariant the kernel relies on does not hold.
//   nvcc -std=c++17 -I iq_tool_b200/csrc -o plan_check tests/native/plan_check.cu   (no GPU needed)
#include <cstdio>
#include "fused_front2.cuh"
using namespace iqgpu;

static int bad = 0;
#define REQUIRE(cond, ...) do { if (!(cond)) { bad++; printf("  VIOLATED: " __VA_ARGS__); printf("\n"); } } while (0)

template <int S> static void check()
{
    using P = W2Plan<S>;
    printf("S=%d warp_f2=%d sup=%d flat_new=%d reg0=%d reg1=%d halo=%lld:", S, P::warp_f2, P::sup, P::flat_new, (int)P::reg0,
           (int)(S >= 2 ? P::reg1 : false), P::halo_frames());
    for (int d = 0; d < S; d++) printf(" [d%d m=%d out=%d R=%d per=%d Hh=%d]", d, P::m(d), P::out(d), P::R(d), P::period(d), P::Hh(d));
    printf("\n");
    REQUIRE(P::warp_f2 % 2 == 0, "per-warp region must keep 16-byte alignment");
    REQUIRE(P::flat_off % 2 == 0, "flat level must be 16-byte aligned");
    REQUIRE(P::flat_size >= W2_ARB_HIST + P::flat_new + 2, "two zero entries beyond the newest flat sample (two-output polyphase lanes)");
    int taps = 0;
    for (int d = 0; d < S; d++) {
        taps += 2 * P::m(d);
        REQUIRE(P::out(d) == 32 * P::R(d) && P::R(d) >= 2, "stage %d: whole register tiles of >= 2 outputs", d);
        REQUIRE(P::Hh(d) % P::R(d) == 0 && P::Hh(d) >= 2 * P::m(d) - 1, "stage %d: history covers the window in whole groups", d);
        REQUIRE(2 * P::Hh(d) <= 64, "stage %d: history slide fits two rounds", d);
        REQUIRE(P::e_off(d) % 2 == 0 && P::o_off(d) % 2 == 0, "stage %d: planes 16-byte aligned", d);
        REQUIRE((P::period(d) & (P::period(d) - 1)) == 0, "stage %d: period is a power of two", d);
        REQUIRE(P::period(d) * P::nat(d) == P::out(d), "stage %d: a run consumes exactly `period` ticks", d);
        if (d + 1 < S) {
            const int r = P::ratio(d);
            REQUIRE(r == 1 || r == 2, "stage %d: consumer takes one or two runs", d);
            REQUIRE(P::period(d + 1) == r * P::period(d), "stage %d: consumer period", d);
            REQUIRE(r == 1 ? (P::R(d) == 2 * P::R(d + 1)) : (P::R(d) == P::R(d + 1)), "stage %d: tile sizes of producer and consumer", d);
        }
    }
    REQUIRE(taps <= W2_MAX_TAPS, "taps fit the parameter block");
    REQUIRE(P::sup == (S ? P::period(S - 1) : 1), "super-tick = period of the last stage");
    if (P::reg0) REQUIRE(P::m(0) == 3 || P::m(0) == 5, "register first stage exists for semi-lengths 3 and 5");
    if (S >= 2 && P::reg1) REQUIRE(P::m(1) == 3 && P::R(1) == 4 && P::m(0) == 3, "register second stage: semi-length 3, 4 outputs per lane");
    // one CTA per SM: bank image + (NCO table) + warps
    const size_t avail = 227 * 1024 - 1024, fixed = (size_t)W2_BANK_F2 * 8 + 1024 * 8;
    const int warps = (int)((avail - fixed) / ((size_t)P::warp_f2 * 8));
    REQUIRE(warps >= 16, "at least 16 warps per CTA fit next to the tables (got %d)", warps);
}

int main()
{
    check<0>(); check<1>(); check<2>(); check<3>(); check<4>(); check<5>(); check<6>();
    REQUIRE(w2_bank_row(255) + 7 <= W2_BANK_F2 - 1, "bank image holds the last row plus one zero float2");
    REQUIRE(w2_bank_row(0) >= 2, "two zero float2 in front of the first row");
    for (int i = 0; i + 1 < 256; i++) REQUIRE(w2_bank_row(i + 1) - w2_bank_row(i) >= 9, "rows %d/%d separated by >= 2 zero float2", i, i + 1);
    printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
