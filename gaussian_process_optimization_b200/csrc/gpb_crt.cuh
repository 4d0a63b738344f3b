// Modular (CRT) splitting for the int8 engine of gpb_ozaki.cu: the integer arithmetic shared by the device kernels and a host
// restatement (gpb_ozaki_crt_host_*, used by the CPU tests to hold this code bit-identical to oracle/ozaki_emulation.py).
//
// An fp64 operand row is scaled by a power of two to a beta-bit integer Q (|Q| < 2^(beta - 1), beta <= 62); for NMOD pairwise
// coprime moduli p_i <= 256 the balanced residues Q mod p_i are int8 planes, ONE exact int8 product per modulus gives
// (A' B'^T) mod p_i, and the integer product X, |X| < P / 2, is rebuilt from its residues by Garner's mixed-radix form
//     X = v_0 + v_1 p_0 + v_2 p_0 p_1 + ...,      v_i = e_i r_i - sum_{j < i} d_ij v_j   (mod p_i, balanced),
// e_i = (p_0 ... p_{i-1})^-1 mod p_i,  d_ij = e_i (p_0 ... p_{j-1}) mod p_i  (compile-time constants, |.| <= 128: the sum stays
// below 2^19 and is reduced ONCE per modulus by an exact multiply-high division), then evaluated in fp64: three mixed-radix
// digits at a time exactly in int32, Horner over those groups with one rounded product and one rounded sum per step.
#pragma once
#include <stdint.h>

#include <utility>

#if defined(__CUDACC__)
#define GPB_HD __host__ __device__ inline
#else
#define GPB_HD inline
#endif

namespace gpb {
namespace crt {

constexpr int MAXMOD = 18, MINMOD = 10;

GPB_HD constexpr int modulus(int i) {
  constexpr int t[MAXMOD] = {256, 255, 253, 251, 247, 241, 239, 233, 229, 227, 223, 217, 211, 199, 197, 193, 191, 181};
  return t[i];
}
GPB_HD constexpr int balanced(long long x, int p) {                      // representative in [-(p / 2), (p - 1) / 2]
  long long r = x % p;
  if (r < 0) r += p;
  return (int)(r >= (p + 1) / 2 ? r - p : r);
}
GPB_HD constexpr int inverse_mod(int a, int p) {                          // brute force; compile time only
  int am = a % p;
  if (am < 0) am += p;
  for (int x = 1; x < p; ++x)
    if ((am * x) % p == 1) return x;
  return 0;
}
GPB_HD constexpr int prefix_mod(int j, int p) {                           // p_0 ... p_{j-1} mod p
  long long r = 1 % p;
  for (int l = 0; l < j; ++l) r = (r * (modulus(l) % p)) % p;
  return (int)r;
}
GPB_HD constexpr int e_coef(int i) { return balanced(inverse_mod(prefix_mod(i, modulus(i)), modulus(i)), modulus(i)); }
GPB_HD constexpr int d_coef(int i, int j) {
  return balanced((long long)inverse_mod(prefix_mod(i, modulus(i)), modulus(i)) * prefix_mod(j, modulus(i)), modulus(i));
}
GPB_HD constexpr uint32_t magic(int p) { return (uint32_t)((1ull << 32) / (unsigned)p) + 1u; }   // floor(u / p) = mulhi(u, magic) for u < 2^24
GPB_HD constexpr int pow2_mod(int bits, int p) {                          // 2^bits mod p, balanced
  long long r = 1 % p;
  for (int b = 0; b < bits; ++b) r = (r * 2) % p;
  return balanced(r, p);
}

// Balanced residue of t in two instructions: tb = t + K p >= 0 (the multiple of p is folded into the sum by the caller, tb + p / 2 < 2^24),
// q = floor((tb + p / 2) / p) as the high word of tb * magic + (p / 2) * magic (one wide multiply-add), and tb - q p is the residue.
GPB_HD constexpr int multiple_for(int p, int lim) { return ((lim + p - 1) / p) * p; }
GPB_HD constexpr uint64_t half_magic(int p) { return (uint64_t)(p / 2) * magic(p); }
GPB_HD int reduce_folded(int tb, int p, uint32_t mg, uint64_t hmg) {
  const uint32_t q = (uint32_t)(((uint64_t)(uint32_t)tb * mg + hmg) >> 32);
  return tb - (int)(q * (uint32_t)p);
}
GPB_HD uint32_t pack4(int r0, int r1, int r2, int r3) {                   // the low bytes of four residues
#if defined(__CUDA_ARCH__)
  return __byte_perm(__byte_perm((uint32_t)r0, (uint32_t)r1, 0x0040), __byte_perm((uint32_t)r2, (uint32_t)r3, 0x0040), 0x5410);
#else
  return ((uint32_t)r0 & 0xFFu) | (((uint32_t)r1 & 0xFFu) << 8) | (((uint32_t)r2 & 0xFFu) << 16) | (((uint32_t)r3 & 0xFFu) << 24);
#endif
}

// ---- per-modulus parameters for the loops that run over the moduli at run time (residue extraction, the drains) -----------
struct ModParams {
  int p;
  uint32_t mg;
  int init_res;      // residue extraction: (a multiple of p >= 8 * 255 * 128 + 128) - (2^62 mod p): the operand is biased by 2^62
  int kp_acc;        // drains: a multiple of p >= |hi * c16 + lo|, <= 2^15 * 128 + 2^16
  int c16;           // 2^16 mod p, balanced
  uint32_t w_lo, w_hi;   // 2^(8 k) mod p, balanced, k = 0..3 / 4..7, packed as signed bytes
  int pad;
  uint64_t hmg;      // (p / 2) * magic
};
constexpr int LIM_RES = 8 * 255 * 128 + 128, LIM_ACC = (1 << 22) + (1 << 16);
GPB_HD constexpr ModParams mod_params(int i) {
  const int p = modulus(i);
  ModParams m = {};
  m.p = p;
  m.mg = magic(p);
  m.init_res = multiple_for(p, LIM_RES) - pow2_mod(62, p);
  m.kp_acc = multiple_for(p, LIM_ACC);
  m.c16 = pow2_mod(16, p);
  m.hmg = half_magic(p);
  uint32_t lo = 0, hi = 0;
  for (int k = 0; k < 4; ++k) {
    lo |= ((uint32_t)pow2_mod(8 * k, p) & 0xFFu) << (8 * k);
    hi |= ((uint32_t)pow2_mod(8 * (k + 4), p) & 0xFFu) << (8 * k);
  }
  m.w_lo = lo;
  m.w_hi = hi;
  return m;
}
struct ModTable { ModParams m[MAXMOD]; };
GPB_HD constexpr ModTable make_table() {
  ModTable t = {};
  for (int i = 0; i < MAXMOD; ++i) t.m[i] = mod_params(i);
  return t;
}

GPB_HD int dot_u8_s8(uint32_t a, uint32_t b, int c) {                      // sum of (unsigned byte of a) * (signed byte of b) + c
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  for (int k = 0; k < 4; ++k) c += (int)((a >> (8 * k)) & 0xFFu) * (int)(int8_t)((b >> (8 * k)) & 0xFFu);
  return c;
#endif
}
// balanced residue of the integer q, |q| <= 2^61, given u = q + 2^62 > 0: the bytes of u weighted by 2^(8 k) mod p, minus 2^62 mod p
// (no sign handling; the constant sits in the accumulator the first dp4a starts from)
GPB_HD uint64_t bias_operand(long long q) { return (uint64_t)q + (1ull << 62); }
GPB_HD int residue_of_biased(uint64_t u, const ModParams &m) {
  int t = dot_u8_s8((uint32_t)u, m.w_lo, m.init_res);
  t = dot_u8_s8((uint32_t)(u >> 32), m.w_hi, t);
  return reduce_folded(t, m.p, m.mg, m.hmg);
}
GPB_HD int residue_of(long long q, const ModParams &m) { return residue_of_biased(bias_operand(q), m); }
// balanced residue of an int32 accumulation s = hi 2^16 + lo
GPB_HD int residue_of_sum(int s, const ModParams &m) {
  const int t = (s >> 16) * m.c16 + ((s & 0xFFFF) + m.kp_acc);
  return reduce_folded(t, m.p, m.mg, m.hmg);
}

// ---- reconstruction (compile-time moduli) -----------------------------------------------------------------------------------
template <int I, int J> struct DCoef { static constexpr int value = d_coef(I, J); };
template <int I> struct ECoef { static constexpr int value = e_coef(I); };
template <int I> struct Mod {
  static constexpr int p = modulus(I);
  static constexpr uint32_t mg = magic(modulus(I));
  static constexpr int kp = multiple_for(modulus(I), (I + 1) * 128 * 128);
  static constexpr uint64_t hmg = half_magic(modulus(I));
};
// -d_ij for j = 4 K .. 4 K + 3 (zero from j = I on) as signed bytes: the Garner sum as dp4a over packed digits (|d| <= 127: p_i odd)
GPB_HD constexpr uint32_t d_pack_neg(int i, int k) {
  uint32_t w = 0;
  for (int b = 0; b < 4; ++b) {
    const int j = 4 * k + b;
    if (j < i) w |= ((uint32_t)(-d_coef(i, j)) & 0xFFu) << (8 * b);
  }
  return w;
}
template <int I, int K> struct DPackNeg { static constexpr uint32_t value = d_pack_neg(I, K); };
GPB_HD int dot_s8_s8(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
  int d;
  asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  for (int k = 0; k < 4; ++k) c += (int)(int8_t)((a >> (8 * k)) & 0xFFu) * (int)(int8_t)((b >> (8 * k)) & 0xFFu);
  return c;
#endif
}
GPB_HD uint32_t insert_byte(uint32_t w, int v, int b) {                   // byte b of w := low byte of v (b is a compile-time constant)
#if defined(__CUDA_ARCH__)
  return __byte_perm(w, (uint32_t)v, b == 0 ? 0x3214 : b == 1 ? 0x3240 : b == 2 ? 0x3410 : 0x4210);
#else
  return (w & ~(0xFFu << (8 * b))) | (((uint32_t)v & 0xFFu) << (8 * b));
#endif
}

template <int I, int... J> GPB_HD int garner_sum(const int *v, std::integer_sequence<int, J...>) {
  return (0 + ... + (v[J] * DCoef<I, J>::value));
}
template <int I, int NMOD> GPB_HD void garner_steps(const int *r, int *v) {
  if constexpr (I == 0) {
    v[0] = r[0];
  } else {
    const int tb = (r[I] * ECoef<I>::value + Mod<I>::kp) - garner_sum<I>(v, std::make_integer_sequence<int, I>{});
    v[I] = reduce_folded(tb, Mod<I>::p, Mod<I>::mg, Mod<I>::hmg);
  }
  if constexpr (I + 1 < NMOD) garner_steps<I + 1, NMOD>(r, v);
}
// the same digits with the sums as dp4a over the digits found so far, packed four to a word (vp: (NMOD + 3) / 4 words, zeroed)
template <int I, int... K> GPB_HD int garner_sum_packed(const uint32_t *vp, int acc, std::integer_sequence<int, K...>) {
  ((acc = dot_s8_s8(vp[K], DPackNeg<I, K>::value, acc)), ...);
  return acc;
}
template <int I, int NMOD> GPB_HD void garner_steps_packed(const int *r, int *v, uint32_t *vp) {
  if constexpr (I == 0) {
    v[0] = r[0];
  } else {
    const int tb = garner_sum_packed<I>(vp, r[I] * ECoef<I>::value + Mod<I>::kp, std::make_integer_sequence<int, (I + 3) / 4>{});
    v[I] = reduce_folded(tb, Mod<I>::p, Mod<I>::mg, Mod<I>::hmg);
  }
  if constexpr (I + 1 < NMOD) {
    vp[I / 4] = insert_byte(vp[I / 4], v[I], I % 4);
    garner_steps_packed<I + 1, NMOD>(r, v, vp);
  }
}

GPB_HD double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  volatile double r = a * b;   // the product is rounded to double before the sum (no contraction on the host either)
  return r;
#endif
}
GPB_HD double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  volatile double r = a + b;
  return r;
#endif
}
// groups of three mixed-radix digits, exact in int32 (< 2^24); G = number of groups
template <int G0, int NMOD> GPB_HD int group_value(const int *v) {
  constexpr int i = 3 * G0;
  int g = v[i];
  if constexpr (i + 2 < NMOD) g += modulus(i) * (v[i + 1] + modulus(i + 1) * v[i + 2]);
  else if constexpr (i + 1 < NMOD) g += modulus(i) * v[i + 1];
  return g;
}
template <int G0> struct GroupWeight {   // product of the three moduli of group G0 (exact in fp64)
  static constexpr double value = (double)modulus(3 * G0) * (double)modulus(3 * G0 + 1) * (double)modulus(3 * G0 + 2);
};
template <int G0, int NMOD> GPB_HD double horner(const int *v, double x) {
  x = add_rn(mul_rn(x, GroupWeight<G0>::value), (double)group_value<G0, NMOD>(v));
  if constexpr (G0 > 0) return horner<G0 - 1, NMOD>(v, x);
  else return x;
}
// the integer with balanced residues r[0 .. NMOD) (|X| < P / 2), rounded to fp64
template <int NMOD, bool PACKED = false> GPB_HD double reconstruct(const int *r) {
  int v[NMOD];
  if constexpr (PACKED) {
    uint32_t vp[(NMOD + 3) / 4] = {};
    garner_steps_packed<0, NMOD>(r, v, vp);
  } else {
    garner_steps<0, NMOD>(r, v);
  }
  constexpr int G = (NMOD + 2) / 3;
  double x = (double)group_value<G - 1, NMOD>(v);
  if constexpr (G > 1) x = horner<G - 2, NMOD>(v, x);
  return x;
}

// bits per operand: k products of two beta-bit integers stay below P / 2 (log2 P from exact integer arithmetic: P is compared
// against powers of two as a 160-bit integer, so that the host, the device and the NumPy restatement agree)
inline int operand_bits(int nmod, long long k) {
  uint32_t w[6] = {1, 0, 0, 0, 0, 0};                                    // P, little endian
  for (int i = 0; i < nmod; ++i) {
    uint64_t carry = 0;
    for (int l = 0; l < 6; ++l) {
      const uint64_t x = (uint64_t)w[l] * (uint32_t)modulus(i) + carry;
      w[l] = (uint32_t)x;
      carry = x >> 32;
    }
  }
  int fl = 0;                                                             // floor(log2 P)
  for (int l = 5; l >= 0; --l)
    if (w[l]) {
      int b = 31;
      while (!((w[l] >> b) & 1u)) --b;
      fl = 32 * l + b;
      break;
    }
  int lk = 0;
  while ((1ll << lk) < k) ++lk;                                           // ceil(log2 k)
  // need k 2^(2 (beta - 1)) <= 2^(lk + 2 beta - 2) <= 2^(fl - 1) <= P / 2
  int beta = (fl + 1 - lk) / 2;
  return beta > 62 ? 62 : beta;
}

}  // namespace crt
}  // namespace gpb
