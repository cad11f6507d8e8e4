/*
 * fft_regs.cuh - in-register complex DFTs of size 2..64 for one thread.
 *
 * Building block of the windowed-FFT kernel (fft_power.cuh) that replaces the
 * reference's fft1D_1024 OpenCL program (lib/fosphor/fft.cl:397-466; its
 * radix-8 / radix-2 local-memory passes are NOT reproduced here).  Each CUDA
 * thread transforms R points held in registers with a fully unrolled
 * decimation-in-frequency recursion whose twiddles are compile-time constants
 * (trivial ones are folded away), so a 1024-point spectrum is two radix-32
 * register passes and ONE shared-memory exchange instead of the reference's
 * four passes and eight barriers.
 *
 * Forward transform, unnormalised:  X[k] = sum_n x[n] exp(-2 pi i n k / R).
 * After dif<R>(v) the spectrum is in bit-reversed register order:
 *   X[k] == v[brev<R>(k)].
 */
#pragma once
#include <cuda_runtime.h>

namespace fosphor_b200 {

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
	if constexpr (I < N) {
		f(std::integral_constant<int, I>{});
		static_for<I + 1, N>(f);
	}
}

__host__ __device__ constexpr int ilog2c(int v) { return v <= 1 ? 0 : 1 + ilog2c(v >> 1); }

/* bit reversal of k within log2(R) bits */
template <int R>
__host__ __device__ constexpr int brev(int k)
{
	int r = 0;
	for (int b = 0; b < ilog2c(R); b++)
		if (k & (1 << b))
			r |= 1 << (ilog2c(R) - 1 - b);
	return r;
}

/* cos / sin of 2 pi q / 64, q = 0..16 (the rest by symmetry) */
template <int Q> struct Cs64;
template <> struct Cs64<0>  { static constexpr float c = 1.0f, s = 0.0f; };
template <> struct Cs64<1>  { static constexpr float c = 0.99518472667219693f, s = 0.098017140329560604f; };
template <> struct Cs64<2>  { static constexpr float c = 0.98078528040323043f, s = 0.19509032201612825f; };
template <> struct Cs64<3>  { static constexpr float c = 0.95694033573220882f, s = 0.29028467725446233f; };
template <> struct Cs64<4>  { static constexpr float c = 0.92387953251128674f, s = 0.38268343236508978f; };
template <> struct Cs64<5>  { static constexpr float c = 0.88192126434835505f, s = 0.47139673682599764f; };
template <> struct Cs64<6>  { static constexpr float c = 0.83146961230254524f, s = 0.55557023301960218f; };
template <> struct Cs64<7>  { static constexpr float c = 0.77301045336273699f, s = 0.63439328416364549f; };
template <> struct Cs64<8>  { static constexpr float c = 0.70710678118654757f, s = 0.70710678118654757f; };
template <> struct Cs64<9>  { static constexpr float c = 0.63439328416364549f, s = 0.77301045336273699f; };
template <> struct Cs64<10> { static constexpr float c = 0.55557023301960218f, s = 0.83146961230254524f; };
template <> struct Cs64<11> { static constexpr float c = 0.47139673682599764f, s = 0.88192126434835505f; };
template <> struct Cs64<12> { static constexpr float c = 0.38268343236508978f, s = 0.92387953251128674f; };
template <> struct Cs64<13> { static constexpr float c = 0.29028467725446233f, s = 0.95694033573220882f; };
template <> struct Cs64<14> { static constexpr float c = 0.19509032201612825f, s = 0.98078528040323043f; };
template <> struct Cs64<15> { static constexpr float c = 0.098017140329560604f, s = 0.99518472667219693f; };
template <> struct Cs64<16> { static constexpr float c = 0.0f, s = 1.0f; };

/* v *= exp(-2 pi i Q / 64), 0 <= Q < 32, with the trivial cases folded */
template <int Q>
__device__ __forceinline__ float2 mul_w64(float2 v)
{
	static_assert(Q >= 0 && Q < 32, "twiddle index");
	if constexpr (Q == 0) {
		return v;
	} else if constexpr (Q == 16) {           /* -i */
		return make_float2(v.y, -v.x);
	} else if constexpr (Q == 8) {            /* (1 - i) / sqrt2 */
		constexpr float h = Cs64<8>::c;
		return make_float2((v.x + v.y) * h, (v.y - v.x) * h);
	} else if constexpr (Q == 24) {           /* (-1 - i) / sqrt2 */
		constexpr float h = Cs64<8>::c;
		return make_float2((v.y - v.x) * h, -(v.x + v.y) * h);
	} else if constexpr (Q < 16) {
		constexpr float c = Cs64<Q>::c, s = Cs64<Q>::s;   /* w = c - i s */
		return make_float2(fmaf(v.y, s, v.x * c), fmaf(-v.x, s, v.y * c));
	} else {
		/* Q in 17..31: w = -cos(pi - a) - i sin(pi - a), a = 2 pi Q / 64 */
		constexpr float c = -Cs64<32 - Q>::c, s = Cs64<32 - Q>::s;
		return make_float2(fmaf(v.y, s, v.x * c), fmaf(-v.x, s, v.y * c));
	}
}

/* Complex add / subtract as ONE packed instruction each (Blackwell FADD2 / FFMA2, add.rn.f32x2 and
 * fma.rn.f32x2 on a 64-bit register pair): the butterflies' 4 scalar FADDs become 2 issue slots.  The
 * kernels are issue bound as soon as the SM clock sags under the board's power cap (sustained runs:
 * ~1.55 GHz), so halving 960 of the N = 1024 kernel's ~2000 instructions is worth more than it looks
 * at burst clocks, where HBM is the roof.  Roundings are those of the scalar form: a + b and
 * fma(b, -1, a) = round(a - b), each component on its own. */
__device__ __forceinline__ float2 cadd(float2 a, float2 b)
{
	return __fadd2_rn(a, b);
}

__device__ __forceinline__ float2 csub(float2 a, float2 b)
{
	return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
}

/* One DIF level of size M on registers v[BASE .. BASE+M), then recurse. */
template <int M, int BASE, int R>
__device__ __forceinline__ void dif_level(float2 (&v)[R])
{
	if constexpr (M >= 2) {
		constexpr int H = M / 2;
		static_for<0, H>([&](auto jc) {
			constexpr int j = decltype(jc)::value;
			const float2 a = v[BASE + j], b = v[BASE + j + H];
			v[BASE + j] = cadd(a, b);
			v[BASE + j + H] = mul_w64<j * (64 / M)>(csub(a, b));
		});
		dif_level<H, BASE, R>(v);
		dif_level<H, BASE + H, R>(v);
	}
}

template <int R>
__device__ __forceinline__ void dif(float2 (&v)[R])
{
	static_assert(R >= 2 && R <= 64 && (R & (R - 1)) == 0, "radix");
	dif_level<R, 0, R>(v);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 w)
{
	return make_float2(fmaf(-a.y, w.y, a.x * w.x), fmaf(a.y, w.x, a.x * w.y));
}

} /* namespace fosphor_b200 */
