// Shared pieces of the prefix-LM attention kernels (flash.cu: mma.sync; flash_tc.cu: tcgen05 / TMEM): argument block, the
// closed-form mask of t2s_model.py:456-479 and the counter-based probability dropout that forward and backward regenerate.
#pragma once
#include "evk_common.cuh"

namespace evk {

struct FlashArgs {
  const float *q, *k, *v;    // [B, L, ld] (+ h*32), usually three column blocks of one in_proj output
  int ld;
  float* o; int ldo;         // [B, L, ldo]
  float* lse;                // [B*H, L]  (log2 domain)
  const float *dout; int lddo;
  const float* delta;        // [B*H, L]
  float *dq, *dk, *dv; int lddq;   // [B, L, lddq] (+ h*32)
  int B, H, L, X;
  const long long *xlen, *ylen;
  float scale, p_drop;
  const unsigned long long* rng; unsigned long long sid;
};

namespace {

constexpr int DK = 32;       // head dim
constexpr float LOG2E = 1.4426950408889634f;


// 2^x on the SFU without ex2()'s range handling (ex2.approx.ftz: 2^-22 relative error, 2^-inf = 0): the kernels evaluate it
// for every score and are issue-bound
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ bool allowed(int i, int j, int X, int xl, int yl) {
  const bool text = j < xl, audio = ((j - X) < yl) & (j <= i);       // both evaluated: selects, no divergent branch
  return (j < X) ? text : audio;
}
// Probability dropout.  One 32-bit hash now decides a 2 x 2 block of (query, key) pairs through four 16-bit fields
// (drop iff field < round(p * 65536); p = 0.1 -> 0.100006), where round 1 hashed every element separately (10..15 integer
// instructions per element in kernels whose useful work is ~20).  Forward and both backward kernels regenerate the same
// mask from (seed, offset, stream id, b, h, i, j); nothing is stored.
struct DropKey { uint32_t s0, s1, thr; float inv; };
__device__ __forceinline__ DropKey drop_key(const FlashArgs& a) {
  DropKey d{0u, 0u, 0u, 1.f};
  if (a.p_drop > 0.f) {
    Philox ph(a.rng[0]);
    uint4 r = ph(a.rng[1], a.sid);
    d.s0 = r.x; d.s1 = r.y;
    d.thr = (uint32_t)fminf(a.p_drop * 65536.f + 0.5f, 65535.f);
    d.inv = 1.f / (1.f - a.p_drop);
  }
  return d;
}
// hash of a PAIR of query rows: index = z * ceil(L / 2) + (i >> 1)
__device__ __forceinline__ uint32_t drop_row(const DropKey& d, uint32_t rowpair) {
  uint32_t x = rowpair * 0x9E3779B1u + d.s0;
  x ^= x >> 15; x *= 0x85EBCA77u; x ^= x >> 13;
  return x;
}
// 32 bits for rows (2r, 2r+1) x columns (2c, 2c+1): even row = the value itself, odd row = one more mixing round
__device__ __forceinline__ uint32_t drop_block(const DropKey& d, uint32_t rowh, uint32_t colpair) {
  uint32_t x = rowh ^ (colpair * 0xC2B2AE3Du + d.s1);
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t drop_odd(uint32_t x) { x *= 0x9E3779B1u; return x ^ (x >> 15); }
// keep flags of (row i, cols j, j+1), j even
__device__ __forceinline__ void drop_pair(const DropKey& d, uint32_t rowh, int i, int j, bool& k0, bool& k1) {
  uint32_t x = drop_block(d, rowh, (uint32_t)j >> 1);
  if (i & 1) x = drop_odd(x);
  k0 = (x & 0xffffu) >= d.thr;
  k1 = (x >> 16) >= d.thr;
}

}  // namespace
}  // namespace evk
