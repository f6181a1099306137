// Packed fp32 pairs for sm_100a: fma / mul / add .rn.f32x2 (SASS FFMA2 — two IEEE fp32 operations per ISSUED instruction).
//
// The per-image tail kernels (convblock_fused.cu) are bound by issued instructions per element, not by bytes or by a
// math pipe (profiles/r2_tail_bwd_lines.md: ~100 thread instructions per element at IPC 1.9 with four warps per
// scheduler), and ~40 % of those instructions are fp32 FMA / MUL / ADD on eight channels that all see the same
// formula — i.e. natural pairs.  Every lane of a pair is the same round-to-nearest IEEE operation as the scalar
// instruction, so results are bit-identical to the scalar formulation as long as the operation ORDER is kept.
//
// Operands are float2 values; ptxas allocates aligned 64-bit register pairs and knows three operand forms
// (cuobjdump: R.F32x2.HI_LO, the swapped R.F32x2.LO_HI, and the scalar broadcast R.F32), so bc2(s) costs nothing.
#pragma once
#include <cuda_runtime.h>

namespace pcm {

__device__ __forceinline__ float2 bc2(float s) { return make_float2(s, s); }

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// three-input maximum (FMNMX3)
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

}  // namespace pcm
