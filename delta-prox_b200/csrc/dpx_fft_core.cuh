// dpx_fft_core.cuh — in-shared-memory power-of-two FFT building blocks for the fused engine.
//
// Compiles under nvcc (device code) and under g++ with -DDPX_EMU (tests/emu: a thread-per-CUDA-thread
// emulator used to verify the index arithmetic on the CPU box, where there is no GPU).
//
// A tile holds COLS independent length-N complex sequences, element (n, c) at float2 index
//     phys(n, c) = (n + (n >> 3)) * COLS + c
// (columns interleaved; one padding point-row after every 8 points, which makes every pass below
// shared-memory bank-conflict free per half-warp for COLS in {2,4} and a radix-8 last pass).
//
// Forward transform = decimation in frequency with radices (RA, RB, RC), N = RA*RB*RC:
//     natural order in  ->  digit-reversed order out:  frequency k = qa + RA*qb + RA*RB*qc  sits at
//     position pos(k) = qa*(N/RA) + qb*RC + qc.
// Inverse transform = the same passes run backwards with conjugated twiddles (decimation in time):
//     digit-reversed in -> natural order out, unnormalised.
// Working in digit-reversed order between the two costs nothing here because everything done in the
// frequency domain (the spectral solve) is element-wise: its constants are simply stored pre-permuted.
#pragma once

#ifdef DPX_EMU
#include "../../tests/emu/cuda_emu.h"
#else
#include <cuda_runtime.h>
#define DPX_HD __device__ __forceinline__
#endif

namespace dpx {
namespace fft {

// Complex add / sub as ONE packed instruction (Blackwell FADD2: add.f32x2 / sub.f32x2 on an aligned register pair, same
// IEEE round-to-nearest results as two scalar FADDs).  The butterflies are add-dominated (44 % of k_col's SASS was FADD) and
// both fused kernels are limited by instruction issue before the FP32 pipes fill, so halving the add instructions is a
// direct win (profiles/README.md, step 6).  -DDPX_NO_F32X2 or the CPU emulator use the scalar form.
#if defined(__CUDA_ARCH__) && !defined(DPX_EMU) && !defined(DPX_NO_F32X2)
DPX_HD float2 cadd(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
DPX_HD float2 csub(float2 a, float2 b) {
  float2 r;
  asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.f32x2 rc, ra, rb; mov.b64 {%0,%1}, rc; }"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
#else
DPX_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
DPX_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
DPX_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
DPX_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a*conj(b)
// multiply by -i (forward) / +i (inverse)
template <bool INV>
DPX_HD float2 mul_mi(float2 a) { return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x); }

// exp(-2 pi i k / 16), k = 0..15 (cos, sin tables; sign handled by the caller)
#define DPX_C16_1 0.92387953251128674f
#define DPX_S16_1 0.38268343236508977f
#define DPX_R2 0.70710678118654752f

template <bool INV>
DPX_HD float2 w16(int k) {   // w16^k forward, conj for inverse; k compile-time after unrolling
  const float c[16] = {1.f, DPX_C16_1, DPX_R2, DPX_S16_1, 0.f, -DPX_S16_1, -DPX_R2, -DPX_C16_1,
                       -1.f, -DPX_C16_1, -DPX_R2, -DPX_S16_1, 0.f, DPX_S16_1, DPX_R2, DPX_C16_1};
  const float s[16] = {0.f, DPX_S16_1, DPX_R2, DPX_C16_1, 1.f, DPX_C16_1, DPX_R2, DPX_S16_1,
                       0.f, -DPX_S16_1, -DPX_R2, -DPX_C16_1, -1.f, -DPX_C16_1, -DPX_R2, -DPX_S16_1};
  return make_float2(c[k & 15], INV ? s[k & 15] : -s[k & 15]);
}

// ---- small DFTs in registers: out[q] = sum_m a[m] w_R^{mq}, natural order in and out ----------------
template <bool INV>
DPX_HD void dft2(float2& a0, float2& a1) {
  const float2 t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

template <bool INV>
DPX_HD void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
  const float2 s13 = cadd(a1, a3), t = csub(a1, a3);
  a0 = cadd(s02, s13);
  a2 = csub(s02, s13);
  // a1 = d02 + w t, a3 = d02 - w t with w = -i (forward) / +i (inverse): w t = (t.y, -t.x) / (-t.y, t.x).  Scalar adds:
  // a packed add would first need the swapped pair in its own registers
  if (INV) {
    a1 = make_float2(d02.x - t.y, d02.y + t.x);
    a3 = make_float2(d02.x + t.y, d02.y - t.x);
  } else {
    a1 = make_float2(d02.x + t.y, d02.y - t.x);
    a3 = make_float2(d02.x - t.y, d02.y + t.x);
  }
}

template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<2, INV> {
  static DPX_HD void run(float2 (&a)[2]) { dft2<INV>(a[0], a[1]); }
};
template <bool INV>
struct Dft<4, INV> {
  static DPX_HD void run(float2 (&a)[4]) { dft4<INV>(a[0], a[1], a[2], a[3]); }
};
// R = A*B with m = B*m1 + m2, q = q1 + A*q2:
//   y[m2][q1] = DFT_A over m1 of a[B*m1+m2];  y *= w_R^{m2 q1};  out[q1 + A*q2] = DFT_B over m2 of y[m2][q1]
template <bool INV>
struct Dft<8, INV> {   // A = 4, B = 2
  static DPX_HD void run(float2 (&a)[8]) {
    dft4<INV>(a[0], a[2], a[4], a[6]);      // m2 = 0 : y[0][q1] in a[2*q1]
    dft4<INV>(a[1], a[3], a[5], a[7]);      // m2 = 1 : y[1][q1] in a[2*q1+1]
    a[3] = cmul(a[3], w16<INV>(2));         // w8^1
    a[7] = cmul(a[7], w16<INV>(6));         // w8^3
    // out[q1 + 4*q2] = y[0][q1] +/- y[1][q1];  the twiddle w8^2 = -i (forward) / +i (inverse) of y[1][2] = a[5] is folded in
    float2 o[8];
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
      if (q1 == 2) continue;
      o[q1] = cadd(a[2 * q1], a[2 * q1 + 1]);
      o[q1 + 4] = csub(a[2 * q1], a[2 * q1 + 1]);
    }
    if (INV) {
      o[2] = make_float2(a[4].x - a[5].y, a[4].y + a[5].x);
      o[6] = make_float2(a[4].x + a[5].y, a[4].y - a[5].x);
    } else {
      o[2] = make_float2(a[4].x + a[5].y, a[4].y - a[5].x);
      o[6] = make_float2(a[4].x - a[5].y, a[4].y + a[5].x);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = o[i];
  }
};
template <bool INV>
struct Dft<16, INV> {  // A = 4, B = 4
  static DPX_HD void run(float2 (&a)[16]) {
#pragma unroll
    for (int m2 = 0; m2 < 4; ++m2) dft4<INV>(a[m2], a[4 + m2], a[8 + m2], a[12 + m2]);   // y[m2][q1] in a[4*q1+m2]
#pragma unroll
    for (int q1 = 1; q1 < 4; ++q1) {
#pragma unroll
      for (int m2 = 1; m2 < 4; ++m2) a[4 * q1 + m2] = cmul(a[4 * q1 + m2], w16<INV>(m2 * q1));
    }
    float2 o[16];
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
      float2 y0 = a[4 * q1], y1 = a[4 * q1 + 1], y2 = a[4 * q1 + 2], y3 = a[4 * q1 + 3];
      dft4<INV>(y0, y1, y2, y3);
      o[q1] = y0; o[q1 + 4] = y1; o[q1 + 8] = y2; o[q1 + 12] = y3;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = o[i];
  }
};

// radix 3 and 12 = 3 x 4: image sides 3 * 2^k (192 ... 3072; the reference's own test image is 768 x 1024)
#define DPX_S3 0.86602540378443865f
template <bool INV>
DPX_HD float2 w12(int k) {   // exp(-2 pi i k / 12) forward, conj for inverse; k compile-time after unrolling
  const float c[12] = {1.f, DPX_S3, 0.5f, 0.f, -0.5f, -DPX_S3, -1.f, -DPX_S3, -0.5f, 0.f, 0.5f, DPX_S3};
  const float s[12] = {0.f, 0.5f, DPX_S3, 1.f, DPX_S3, 0.5f, 0.f, -0.5f, -DPX_S3, -1.f, -DPX_S3, -0.5f};
  return make_float2(c[k % 12], INV ? s[k % 12] : -s[k % 12]);
}
template <bool INV>
DPX_HD void dft3(float2& a0, float2& a1, float2& a2) {
  const float2 sm = cadd(a1, a2), df = csub(a1, a2);
  const float2 t = make_float2(a0.x - 0.5f * sm.x, a0.y - 0.5f * sm.y);
  const float2 r = make_float2(DPX_S3 * df.x, DPX_S3 * df.y);
  a0 = cadd(a0, sm);
  // forward: y1 = t - i r, y2 = t + i r; inverse: swapped
  const float2 p = make_float2(t.x + r.y, t.y - r.x), q = make_float2(t.x - r.y, t.y + r.x);
  a1 = INV ? q : p;
  a2 = INV ? p : q;
}
template <bool INV>
struct Dft<12, INV> {  // A = 3, B = 4: m = 4 m1 + m2, q = q1 + 3 q2
  static DPX_HD void run(float2 (&a)[12]) {
#pragma unroll
    for (int m2 = 0; m2 < 4; ++m2) dft3<INV>(a[m2], a[4 + m2], a[8 + m2]);            // y[m2][q1] in a[4*q1+m2]
#pragma unroll
    for (int q1 = 1; q1 < 3; ++q1) {
#pragma unroll
      for (int m2 = 1; m2 < 4; ++m2) a[4 * q1 + m2] = cmul(a[4 * q1 + m2], w12<INV>(m2 * q1));
    }
    float2 o[12];
#pragma unroll
    for (int q1 = 0; q1 < 3; ++q1) {
      float2 y0 = a[4 * q1], y1 = a[4 * q1 + 1], y2 = a[4 * q1 + 2], y3 = a[4 * q1 + 3];
      dft4<INV>(y0, y1, y2, y3);
      o[q1] = y0; o[q1 + 3] = y1; o[q1 + 6] = y2; o[q1 + 9] = y3;
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) a[i] = o[i];
  }
};

// radix 5, 10 = 5 x 2 and 20 = 5 x 4: image sides 5 * 2^k (320 ... 2560)
template <bool INV>
DPX_HD float2 w20(int k) {   // exp(-2 pi i k / 20) forward, conj for inverse; k compile-time after unrolling
  const float c[20] = {1.f, 0.95105651629515357f, 0.80901699437494742f, 0.58778525229247313f, 0.30901699437494742f, 0.f,
                       -0.30901699437494742f, -0.58778525229247313f, -0.80901699437494742f, -0.95105651629515357f, -1.f,
                       -0.95105651629515357f, -0.80901699437494742f, -0.58778525229247313f, -0.30901699437494742f, 0.f,
                       0.30901699437494742f, 0.58778525229247313f, 0.80901699437494742f, 0.95105651629515357f};
  const float s[20] = {0.f, 0.30901699437494742f, 0.58778525229247313f, 0.80901699437494742f, 0.95105651629515357f, 1.f,
                       0.95105651629515357f, 0.80901699437494742f, 0.58778525229247313f, 0.30901699437494742f, 0.f,
                       -0.30901699437494742f, -0.58778525229247313f, -0.80901699437494742f, -0.95105651629515357f, -1.f,
                       -0.95105651629515357f, -0.80901699437494742f, -0.58778525229247313f, -0.30901699437494742f};
  return make_float2(c[k % 20], INV ? s[k % 20] : -s[k % 20]);
}
template <bool INV>
DPX_HD void dft5(float2& a0, float2& a1, float2& a2, float2& a3, float2& a4) {
  const float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f, S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;
  const float2 t1 = cadd(a1, a4), t2 = cadd(a2, a3), t3 = csub(a1, a4), t4 = csub(a2, a3);
  const float2 p1 = make_float2(a0.x + C1 * t1.x + C2 * t2.x, a0.y + C1 * t1.y + C2 * t2.y);
  const float2 p2 = make_float2(a0.x + C2 * t1.x + C1 * t2.x, a0.y + C2 * t1.y + C1 * t2.y);
  const float2 q1 = make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y);
  const float2 q2 = make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y);
  a0 = cadd(a0, cadd(t1, t2));
  // forward: y1 = p1 - i q1, y4 = p1 + i q1, y2 = p2 - i q2, y3 = p2 + i q2 (-i q = (q.y, -q.x)); inverse: conjugate roles
  const float2 m1 = make_float2(p1.x + q1.y, p1.y - q1.x), n1 = make_float2(p1.x - q1.y, p1.y + q1.x);
  const float2 m2 = make_float2(p2.x + q2.y, p2.y - q2.x), n2 = make_float2(p2.x - q2.y, p2.y + q2.x);
  a1 = INV ? n1 : m1; a4 = INV ? m1 : n1;
  a2 = INV ? n2 : m2; a3 = INV ? m2 : n2;
}
template <bool INV>
struct Dft<10, INV> {  // A = 5, B = 2: m = 2 m1 + m2, q = q1 + 5 q2
  static DPX_HD void run(float2 (&a)[10]) {
#pragma unroll
    for (int m2 = 0; m2 < 2; ++m2) dft5<INV>(a[m2], a[2 + m2], a[4 + m2], a[6 + m2], a[8 + m2]);      // y[m2][q1] in a[2*q1+m2]
#pragma unroll
    for (int q1 = 1; q1 < 5; ++q1) a[2 * q1 + 1] = cmul(a[2 * q1 + 1], w20<INV>(2 * q1));
    float2 o[10];
#pragma unroll
    for (int q1 = 0; q1 < 5; ++q1) {
      o[q1] = cadd(a[2 * q1], a[2 * q1 + 1]);
      o[q1 + 5] = csub(a[2 * q1], a[2 * q1 + 1]);
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) a[i] = o[i];
  }
};
template <bool INV>
struct Dft<20, INV> {  // A = 5, B = 4: m = 4 m1 + m2, q = q1 + 5 q2
  static DPX_HD void run(float2 (&a)[20]) {
#pragma unroll
    for (int m2 = 0; m2 < 4; ++m2) dft5<INV>(a[m2], a[4 + m2], a[8 + m2], a[12 + m2], a[16 + m2]);    // y[m2][q1] in a[4*q1+m2]
#pragma unroll
    for (int q1 = 1; q1 < 5; ++q1) {
#pragma unroll
      for (int m2 = 1; m2 < 4; ++m2) a[4 * q1 + m2] = cmul(a[4 * q1 + m2], w20<INV>(m2 * q1));
    }
    float2 o[20];
#pragma unroll
    for (int q1 = 0; q1 < 5; ++q1) {
      float2 y0 = a[4 * q1], y1 = a[4 * q1 + 1], y2 = a[4 * q1 + 2], y3 = a[4 * q1 + 3];
      dft4<INV>(y0, y1, y2, y3);
      o[q1] = y0; o[q1 + 5] = y1; o[q1 + 10] = y2; o[q1 + 15] = y3;
    }
#pragma unroll
    for (int i = 0; i < 20; ++i) a[i] = o[i];
  }
};

// radix 9 = 3 x 3 and 15 = 3 x 5: column lengths 1080 = 15*9*8, 720 = 10*9*8, 1440 = 15*12*8, 1200 = 15*10*8, 2160 = 15*9*16 (the
// camera formats whose WIDTH -- 1920, 1280, 2560, 1600, 3840 -- already factors into 2, 3*4, 5*2 and 5*4)
template <bool INV>
DPX_HD float2 w9(int k) {
  const float c[9] = {1.f, 0.76604444311897801f, 0.17364817766693041f, -0.49999999999999978f, -0.93969262078590832f, -0.93969262078590843f, -0.50000000000000044f, 0.17364817766692997f, 0.76604444311897779f};
  const float s[9] = {0.f, 0.64278760968653925f, 0.98480775301220802f, 0.86602540378443871f, 0.34202014332566888f, -0.34202014332566866f, -0.86602540378443837f, -0.98480775301220813f, -0.64278760968653958f};
  return make_float2(c[k % 9], INV ? s[k % 9] : -s[k % 9]);
}
template <bool INV>
DPX_HD float2 w15(int k) {
  const float c[15] = {1.f, 0.91354545764260087f, 0.66913060635885824f, 0.30901699437494745f, -0.10452846326765333f, -0.49999999999999978f, -0.80901699437494734f, -0.97814760073380569f, -0.97814760073380569f, -0.80901699437494756f, -0.50000000000000044f, -0.10452846326765423f, 0.30901699437494723f, 0.66913060635885846f, 0.91354545764260098f};
  const float s[15] = {0.f, 0.40673664307580015f, 0.74314482547739413f, 0.95105651629515353f, 0.9945218953682734f, 0.86602540378443871f, 0.58778525229247325f, 0.20791169081775931f, -0.20791169081775907f, -0.58778525229247303f, -0.86602540378443837f, -0.99452189536827329f, -0.95105651629515364f, -0.74314482547739402f, -0.40673664307580015f};
  return make_float2(c[k % 15], INV ? s[k % 15] : -s[k % 15]);
}
template <bool INV>
struct Dft<9, INV> {   // A = 3, B = 3: m = 3 m1 + m2, q = q1 + 3 q2
  static DPX_HD void run(float2 (&a)[9]) {
#pragma unroll
    for (int m2 = 0; m2 < 3; ++m2) dft3<INV>(a[m2], a[3 + m2], a[6 + m2]);             // y[m2][q1] in a[3*q1+m2]
#pragma unroll
    for (int q1 = 1; q1 < 3; ++q1) {
#pragma unroll
      for (int m2 = 1; m2 < 3; ++m2) a[3 * q1 + m2] = cmul(a[3 * q1 + m2], w9<INV>(m2 * q1));
    }
    float2 o[9];
#pragma unroll
    for (int q1 = 0; q1 < 3; ++q1) {
      float2 y0 = a[3 * q1], y1 = a[3 * q1 + 1], y2 = a[3 * q1 + 2];
      dft3<INV>(y0, y1, y2);
      o[q1] = y0; o[q1 + 3] = y1; o[q1 + 6] = y2;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = o[i];
  }
};
template <bool INV>
struct Dft<15, INV> {  // A = 3, B = 5: m = 5 m1 + m2, q = q1 + 3 q2
  static DPX_HD void run(float2 (&a)[15]) {
#pragma unroll
    for (int m2 = 0; m2 < 5; ++m2) dft3<INV>(a[m2], a[5 + m2], a[10 + m2]);            // y[m2][q1] in a[5*q1+m2]
#pragma unroll
    for (int q1 = 1; q1 < 3; ++q1) {
#pragma unroll
      for (int m2 = 1; m2 < 5; ++m2) a[5 * q1 + m2] = cmul(a[5 * q1 + m2], w15<INV>(m2 * q1));
    }
    float2 o[15];
#pragma unroll
    for (int q1 = 0; q1 < 3; ++q1) {
      float2 y0 = a[5 * q1], y1 = a[5 * q1 + 1], y2 = a[5 * q1 + 2], y3 = a[5 * q1 + 3], y4 = a[5 * q1 + 4];
      dft5<INV>(y0, y1, y2, y3, y4);
      o[q1] = y0; o[q1 + 3] = y1; o[q1 + 6] = y2; o[q1 + 9] = y3; o[q1 + 12] = y4;
    }
#pragma unroll
    for (int i = 0; i < 15; ++i) a[i] = o[i];
  }
};

// ---- tile geometry ------------------------------------------------------------------------------------
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }

template <int N_, int RA_, int RB_, int RC_, int COLS_>
struct Tile {
  static constexpr int N = N_, RA = RA_, RB = RB_, RC = RC_, COLS = COLS_;
  static_assert(RA_ * RB_ * RC_ == N_, "N must equal RA*RB*RC");
  static constexpr int MA = N / RA;          // stride of pass A
  static constexpr int MB = MA / RB;         // stride of pass B  (== RC)
  static constexpr int LOG_MA = ilog2(MA);
  static_assert(MA % 8 == 0, "pass-A stride must be a multiple of 8");
  // padded point index: one spare point after every 8 points and one more after every MA points.  With the
  // thread->task maps used below this makes every pass AND the digit-reversal scatter/gather of the row kernel
  // free of shared-memory bank conflicts per half-warp (COLS in {2,4}).
  static constexpr bool MA_POW2 = (MA & (MA - 1)) == 0;   // odd factors in the later passes (720 = 10*9*8, 1920 = 20*12*8, ...) make MA = 72, 96, ...
  static DPX_HD int pn(int n) { return n + (n >> 3) + (MA_POW2 ? (n >> LOG_MA) : n / MA); }
  static constexpr int PADDED_N = N + N / 8 + RA;
  static constexpr int SMEM_FLOAT2 = PADDED_N * COLS;
  static DPX_HD int phys(int n, int c) { return pn(n) * COLS + c; }
  // offset (in points) of element m of a butterfly whose elements are M apart, relative to its first element;
  // valid for M == MA, for M a multiple of 8 inside one MA-block, and for M == 1 on an RC-aligned block
  template <int M>
  static constexpr bool linear() { return M == MA || M % 8 == 0 || M == 1; }
  template <int M>
  static DPX_HD int delta(int m) {
    return M == MA ? m * (MA + MA / 8 + 1) : (M == 1 ? m + (m >> 3) : m * (M + M / 8));
  }
  // position of frequency k after the forward (DIF) transform
  static DPX_HD int pos_of_freq(int k) {
    const int qa = k % RA, qb = (k / RA) % RB, qc = k / (RA * RB);
    return qa * MA + qb * MB + qc;
  }
  static DPX_HD int freq_of_pos(int p) {
    const int qa = p / MA, qb = (p % MA) / MB, qc = p % MB;
    return qa + RA * qb + RA * RB * qc;
  }
};

// Twiddle records: for a pass of radix R over sub-blocks of length L (M = L/R butterflies j = 0..M-1), the factors
// exp(-2 pi i j q / L), q = 0..R-1 (q = 0 is 1) are stored TRANSPOSED as float4 pairs, rec4[(q/2)*M + j] = (w^{j q}, w^{j (q+1)}):
// the lanes of a warp work on consecutive j, so each of a thread's R/2 128-bit fetches reads one contiguous run of the
// table -- one L1 wavefront per fetch instead of one per distinct j ([j][q] records cost 8 wavefronts per fetch, as much
// LSU time as the data itself; profiles/README.md v6).  (tables are built on the host in double precision)
template <int R, int M>
DPX_HD void load_twiddles(const float2* __restrict__ rec, int j, float2 (&w)[R]) {
  const float4* r4 = reinterpret_cast<const float4*>(rec) + j;
#pragma unroll
  for (int q = 0; q < (R + 1) / 2; ++q) {                 // odd radix: the last record's second factor is padding
    const float4 v = r4[q * M];
    w[2 * q] = make_float2(v.x, v.y);
    if (2 * q + 1 < R) w[2 * q + 1] = make_float2(v.z, v.w);
  }
}

// One shared-memory pass of radix R over sub-blocks of length L (stride M = L/R) for all COLS columns.
//   forward (INV=false): gather, DFT_R, multiply output q by w_L^{j q}, scatter back (same places).
//   inverse (INV=true) : gather, multiply input q by conj(w_L^{j q}), inverse DFT_R, scatter back.
// `rec` = twiddle records of this pass (unused when !TWIDDLE).  All threads of the CTA must call this; the
// caller places __syncthreads() between passes.
template <class T, int R, int L, bool INV, bool TWIDDLE>
DPX_HD void smem_pass(float2* sm, const float2* __restrict__ rec, int tid, int nthreads) {
  constexpr int M = L / R;
  constexpr int NTASK = T::COLS * T::N / R;
  constexpr bool LIN = T::template linear<M>();
  for (int task = tid; task < NTASK; task += nthreads) {
    const int c = task % T::COLS;
    const int t2 = task / T::COLS;
    const int j = t2 % M;
    const int base = (t2 / M) * L + j;
    const int p0 = T::phys(base, c);
    float2 a[R];
#pragma unroll
    for (int m = 0; m < R; ++m) a[m] = sm[LIN ? p0 + T::template delta<M>(m) * T::COLS : T::phys(base + m * M, c)];
    if (TWIDDLE) {
      float2 w[R];
      load_twiddles<R, M>(rec, j, w);
      if (INV) {
#pragma unroll
        for (int q = 1; q < R; ++q) a[q] = cmulc(a[q], w[q]);
        Dft<R, true>::run(a);
      } else {
        Dft<R, false>::run(a);
#pragma unroll
        for (int q = 1; q < R; ++q) a[q] = cmul(a[q], w[q]);
      }
    } else {
      Dft<R, INV>::run(a);
    }
#pragma unroll
    for (int m = 0; m < R; ++m) sm[LIN ? p0 + T::template delta<M>(m) * T::COLS : T::phys(base + m * M, c)] = a[m];
  }
}

// twiddle-record table of a tile: [pass A: MA*RA float2][pass B: MB*RB float2]
template <class T>
struct TwiddleLayout {
  static constexpr int RAE = (T::RA + 1) / 2 * 2, RBE = (T::RB + 1) / 2 * 2;      // factor pairs: odd radices are padded
  static constexpr int A_OFF = 0, B_OFF = T::MA * RAE, TOTAL = T::MA * RAE + T::MB * RBE;
};

// Whole transforms on a tile resident in shared memory.
template <class T>
DPX_HD void tile_fft_forward(float2* sm, const float2* __restrict__ tw, int tid, int nthreads) {
  smem_pass<T, T::RA, T::N, false, true>(sm, tw + TwiddleLayout<T>::A_OFF, tid, nthreads);
  __syncthreads();
  smem_pass<T, T::RB, T::MA, false, true>(sm, tw + TwiddleLayout<T>::B_OFF, tid, nthreads);
  __syncthreads();
  smem_pass<T, T::RC, T::MB, false, false>(sm, tw, tid, nthreads);
  __syncthreads();
}
template <class T>
DPX_HD void tile_fft_inverse(float2* sm, const float2* __restrict__ tw, int tid, int nthreads) {
  smem_pass<T, T::RC, T::MB, true, false>(sm, tw, tid, nthreads);
  __syncthreads();
  smem_pass<T, T::RB, T::MA, true, true>(sm, tw + TwiddleLayout<T>::B_OFF, tid, nthreads);
  __syncthreads();
  smem_pass<T, T::RA, T::N, true, true>(sm, tw + TwiddleLayout<T>::A_OFF, tid, nthreads);
  __syncthreads();
}

}  // namespace fft
}  // namespace dpx
