/* tools/check_sincosf.c -- host check for orbx_glibc_sincosf (extractorb_b200/csrc/orbx_kernels.cuh): the same statements in
 * plain C (double arithmetic, fma), compared with this machine's libm sinf / cosf on EVERY float in [0, 6.5] -- the angles
 * computeOrbDescriptor can see are kpt.angle * (float)(CV_PI/180.f) with kpt.angle in [0, 360] (reference ORBextractor.cc:110).
 *     gcc -O2 -ffp-contract=off -mfma -o /tmp/check_sincosf tools/check_sincosf.c -lm && /tmp/check_sincosf
 * glibc 2.39 (x86-64, both its FMA and non-FMA builds of the routine): 0 mismatches in 1 087 373 313 inputs. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static float poly(double x, double x2, int n, int neg_table) {
    const double sg = neg_table ? -1.0 : 1.0;
    if ((n & 1) == 0) {
        const double x3 = x * x2;
        const double s1 = fma(x2, -0x1.994eb3774cf24p-13, 0x1.1107605230bc4p-7);
        const double x7 = x3 * x2;
        const double s = fma(x3, -0x1.555545995a603p-3, x);
        return (float)fma(x7, s1, s);
    }
    const double x4 = x2 * x2;
    const double c2 = fma(x2, sg * 0x1.99343027bf8c3p-16, sg * -0x1.6c087e89a359dp-10);
    const double c1 = fma(x2, sg * -0x1.ffffffd0c621cp-2, sg * 0x1p0);
    const double x6 = x4 * x2;
    const double c = fma(x4, sg * 0x1.55553e1068f19p-5, c1);
    return (float)fma(x6, c2, c);
}

static void sincosf_restated(float y, float* sinp, float* cosp) {
    const double x = (double)y;
    uint32_t u;
    memcpy(&u, &y, 4);
    const unsigned top = u >> 20;
    if (top < (0x3f490fdbu >> 20)) {
        if (top < (0x39800000u >> 20)) { *sinp = y; *cosp = 1.0f; return; }
        const double x2 = x * x;
        *sinp = poly(x, x2, 0, 0);
        *cosp = poly(x, x2, 1, 0);
        return;
    }
    const double r = x * 0x1.45F306DC9C883p+23;
    const int n = ((int)r + 0x800000) >> 24;
    const double xr = fma(-(double)n, 0x1.921FB54442D18p0, x);
    const double x2 = xr * xr;
    const double ss = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
    const double cs = (((n + 1) & 3) == 1 || ((n + 1) & 3) == 2) ? -1.0 : 1.0;
    *sinp = poly(xr * ss, x2, n, (n & 2) != 0);
    *cosp = poly(xr * cs, x2, n ^ 1, ((n + 1) & 2) != 0);
}

int main(void) {
    long bad = 0, tot = 0;
    const float lim = 6.5f;
    uint32_t ul;
    memcpy(&ul, &lim, 4);
    for (uint32_t u = 0; u <= ul; ++u) {
        float y, s, c;
        memcpy(&y, &u, 4);
        sincosf_restated(y, &s, &c);
        const float rs = sinf(y), rc = cosf(y);
        if (memcmp(&s, &rs, 4) || memcmp(&c, &rc, 4)) {
            if (bad < 5) printf("y=%a: sin %a vs %a, cos %a vs %a\n", y, s, rs, c, rc);
            ++bad;
        }
        ++tot;
    }
    printf("%ld floats in [0, 6.5]: %ld mismatches against libm sinf/cosf\n", tot, bad);
    return bad != 0;
}
