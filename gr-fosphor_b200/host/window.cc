/*
 * window.cc - FFT window generator for the drop-in.
 *
 * The reference sink builds its window with GNU Radio,
 *     gr::fft::window::build(d_fft_window, 1024, 6.76)      lib/base_sink_c_impl.cc:251-255
 * for the seven types its GRC block offers (grc/fosphor_glfw_sink_c.block.yml:5-11).
 * gr-fft (>= 3.9, CMakeLists.txt:34) is a third-party dependency that is not
 * vendored in the reference tree and not installed here (parity unpinned by
 * the reference: it has no window tests, and gr-fft cannot be run here); this
 * restates the published definitions gr-fft implements:
 *   - CONVENTION: every window is SYMMETRIC over ntaps - 1 intervals
 *     (w[i] = f(i / (ntaps - 1)), w[0] == w[ntaps-1]), not the periodic /
 *     "fftbins" form.  (The reference's own DEFAULT window, fosphor.c:108-121,
 *     is the periodic Hamming x 1.855 - that one is fosphor_cu_default_window.)
 *   - generalised cosine sums  w[i] = sum_k (-1)^k c_k cos(2 pi k i / (ntaps-1))
 *     with the coefficient sets named at each case below;
 *   - flat-top = the 5-term HP / SRS set {1, 1.93, 1.29, 0.388, 0.028} / 4.63867
 *     (peak 4.636 / 4.63867 = 0.99942; NOT the ISO 18431-2 / Matlab / scipy set
 *     0.21557895, 0.41663158, ... - the two differ by up to 1.6e-3);
 *   - Kaiser through the zeroth-order modified Bessel function I0.
 * tests/test_window.py pins every type against these definitions evaluated in
 * double, and against scipy's symmetric windows where the sets coincide.  The
 * engine itself takes the window as a plain float array, so parity of the hot
 * path never depends on this file.
 */
#include <cmath>

extern "C" {

/* numbering of gr::fft::window::win_type */
enum fosphor_window_type {
	FOSPHOR_WIN_HAMMING = 0,
	FOSPHOR_WIN_HANN = 1,
	FOSPHOR_WIN_BLACKMAN = 2,
	FOSPHOR_WIN_RECTANGULAR = 3,
	FOSPHOR_WIN_KAISER = 4,
	FOSPHOR_WIN_BLACKMAN_HARRIS = 5,
	FOSPHOR_WIN_BARTLETT = 6,
	FOSPHOR_WIN_FLATTOP = 7,
};

static double bessel_i0(double x)
{
	/* power series of I0: sum ((x/2)^k / k!)^2 */
	double sum = 1.0, term = 1.0;
	const double q = x * x / 4.0;
	for (int k = 1; k < 200; k++) {
		term *= q / ((double)k * (double)k);
		sum += term;
		if (term < 1e-18 * sum)
			break;
	}
	return sum;
}

static void cos_sum(float *w, int n, const double *c, int nc)
{
	const double m = (double)(n - 1);
	for (int i = 0; i < n; i++) {
		double v = 0.0, sign = 1.0;
		for (int k = 0; k < nc; k++) {
			v += sign * c[k] * cos(2.0 * M_PI * (double)k * (double)i / m);
			sign = -sign;
		}
		w[i] = (float)v;
	}
}

/* 0 on success, -1 for an unknown type or n < 2.  beta is used by Kaiser only. */
int fosphor_window_build(int type, int n, double beta, float *w)
{
	if (n < 2 || !w)
		return -1;
	switch (type) {
	case FOSPHOR_WIN_HAMMING: {
		const double c[] = {0.54, 0.46};
		cos_sum(w, n, c, 2);
		return 0;
	}
	case FOSPHOR_WIN_HANN: {
		const double c[] = {0.5, 0.5};
		cos_sum(w, n, c, 2);
		return 0;
	}
	case FOSPHOR_WIN_BLACKMAN: {
		const double c[] = {0.42, 0.5, 0.08};
		cos_sum(w, n, c, 3);
		return 0;
	}
	case FOSPHOR_WIN_BLACKMAN_HARRIS: {      /* 92 dB, 4-term */
		const double c[] = {0.35875, 0.48829, 0.14128, 0.01168};
		cos_sum(w, n, c, 4);
		return 0;
	}
	case FOSPHOR_WIN_FLATTOP: {
		const double s = 4.63867;
		const double c[] = {1.0 / s, 1.93 / s, 1.29 / s, 0.388 / s, 0.028 / s};
		cos_sum(w, n, c, 5);
		return 0;
	}
	case FOSPHOR_WIN_RECTANGULAR:
		for (int i = 0; i < n; i++)
			w[i] = 1.0f;
		return 0;
	case FOSPHOR_WIN_BARTLETT: {
		const double m = (double)(n - 1);
		for (int i = 0; i < n; i++)
			w[i] = (float)(1.0 - fabs(2.0 * (double)i / m - 1.0));
		return 0;
	}
	case FOSPHOR_WIN_KAISER: {
		const double m = (double)(n - 1), ib = 1.0 / bessel_i0(beta);
		for (int i = 0; i < n; i++) {
			const double t = 2.0 * (double)i / m - 1.0;
			w[i] = (float)(bessel_i0(beta * sqrt(fmax(0.0, 1.0 - t * t))) * ib);
		}
		return 0;
	}
	default:
		return -1;
	}
}

} /* extern "C" */
