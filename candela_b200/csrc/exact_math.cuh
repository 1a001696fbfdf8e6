// Arithmetic rules for bit-exact parity with the reference's traversal
// (SURVEY.md §7.3): no FMA contraction, IEEE division, GLSL min/max semantics.
//
// Every float operation on the parity path goes through these wrappers, which
// map to round-to-nearest intrinsics the compiler never contracts.  The
// library is additionally compiled with -fmad=false -prec-div=true -ftz=false.
#pragma once
#include <cuda_runtime.h>

namespace cndl {

struct V3 { float x, y, z; };

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// min / max exactly as the reference's vendored glm 0.9.8.5 evaluates them (glm/detail/func_common.inl:15-28:
// x < y ? x : y and x > y ? x : y) — the forms its CPU code (Physics.cpp) runs, and the ones its GLSL runs when the
// shader files are compiled against glm (GLSL 4.50 §8.3 leaves the result undefined when an operand is NaN).  They
// differ from fminf/fmaxf only when an operand is NaN (or in the sign of a zero result).
__device__ __forceinline__ float glsl_min(float x, float y) { return x < y ? x : y; }
__device__ __forceinline__ float glsl_max(float x, float y) { return x > y ? x : y; }

__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return {fsub(a.x, b.x), fsub(a.y, b.y), fsub(a.z, b.z)}; }
__device__ __forceinline__ V3 vneg(V3 a) { return {-a.x, -a.y, -a.z}; }
// dot: ((ax*bx + ay*by) + az*bz), products rounded separately
__device__ __forceinline__ float vdot(V3 a, V3 b) {
    return fadd(fadd(fmul(a.x, b.x), fmul(a.y, b.y)), fmul(a.z, b.z));
}
// GLSL cross(a,b) = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
    return {fsub(fmul(a.y, b.z), fmul(b.y, a.z)), fsub(fmul(a.z, b.x), fmul(b.z, a.x)), fsub(fmul(a.x, b.y), fmul(b.x, a.y))};
}

// vec3(M * vec4(v, w)) for a column-major mat4, summed like glm 0.9.8.5's operator*
// (type_mat4x4.inl:526-540): (c0*x + c1*y) + (c2*z + c3*w).
__device__ __forceinline__ V3 xform(const float* __restrict__ m, V3 v, float w) {
    V3 r;
    r.x = fadd(fadd(fmul(m[0], v.x), fmul(m[4], v.y)), fadd(fmul(m[8], v.z), fmul(m[12], w)));
    r.y = fadd(fadd(fmul(m[1], v.x), fmul(m[5], v.y)), fadd(fmul(m[9], v.z), fmul(m[13], w)));
    r.z = fadd(fadd(fmul(m[2], v.x), fmul(m[6], v.y)), fadd(fmul(m[10], v.z), fmul(m[14], w)));
    return r;
}

__device__ __forceinline__ bool finite3(V3 a) { return isfinite(a.x) && isfinite(a.y) && isfinite(a.z); }

}  // namespace cndl
