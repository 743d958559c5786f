// obe_models.cuh -- built-in device functors for model_function (NVRTC-safe, no #include).
//
// The reference's model contract (obe_base.py:50-66) is a Python callable
// model(sets, pars, cons) that must broadcast in two orientations.  A Python callable cannot
// run inside a kernel, so each demo model becomes a scalar functor
//     eval(const double* s, const double* p, const double* c, double* y)
// instantiated into the update kernel (particles vary) and the utility kernel (settings vary).
// Rational models use the non-contracting obe_* ops in numpy's operation order, so their
// values are bit-identical to the reference's.
#ifndef OBE_MODELS_CUH
#define OBE_MODELS_CUH

// b + a / (((x - x0) / d)**2 + 1)      demos/find_peak/sequentialLorentzian.py:66-75
struct ObeLorentzianHWHM {
    enum { NS = 1, NP = 3, NCONS = 1, NCH = 1 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double* c, double* y) {
        const double q = obe_div(obe_sub(s[0], p[0]), c[0]);
        y[0] = obe_add(p[2], obe_div(p[1], obe_add(obe_mul(q, q), 1.0)));
    }
};

// update-pass variant: the reciprocal of the constant linewidth is loop-invariant, so the
// particle loop keeps a single division (the utility pass keeps the exact functor above).
template <>
struct ObeUpdateEval<ObeLorentzianHWHM> {
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double* c, double* y) {
        const double q = (s[0] - p[0]) * (1.0 / c[0]);
        y[0] = fma(p[1], obe_rcp_fast(fma(q, q, 1.0)), p[2]);      // denominator >= 1: no special cases
    }
};

// a / ((2 * (x - x0) / d)**2 + 1) + b   demos/numba/numbaLorentzian.py:104
struct ObeLorentzianFWHM {
    enum { NS = 1, NP = 3, NCONS = 1, NCH = 1 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double* c, double* y) {
        const double q = obe_div(obe_mul(2.0, obe_sub(s[0], p[0])), c[0]);
        y[0] = obe_add(obe_div(p[1], obe_add(obe_mul(q, q), 1.0)), p[2]);
    }
};

template <>
struct ObeUpdateEval<ObeLorentzianFWHM> {
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double* c, double* y) {
        const double q = (2.0 * (s[0] - p[0])) * (1.0 / c[0]);
        y[0] = fma(p[1], obe_rcp_fast(fma(q, q, 1.0)), p[2]);
    }
};

// b + a / (((x - x0) * 2 / d)**2 + 1), linewidth is parameter 3   demos/find_peak/seqLor_pdfevolve.py:28
struct ObeLorentzian4P {
    enum { NS = 1, NP = 4, NCONS = 0, NCH = 1 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double*, double* y) {
        const double q = obe_div(obe_mul(obe_sub(s[0], p[0]), 2.0), p[3]);
        y[0] = obe_add(p[2], obe_div(p[1], obe_add(obe_mul(q, q), 1.0)));
    }
};

// 1 - A / (((f - f0) * 2 / lw)**2 + 1)   demos/server/server_script.py:33
struct ObeLorentzianDip {
    enum { NS = 1, NP = 3, NCONS = 0, NCH = 1 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double*, double* y) {
        const double q = obe_div(obe_mul(obe_sub(s[0], p[0]), 2.0), p[2]);
        y[0] = obe_sub(1.0, obe_div(p[1], obe_add(obe_mul(q, q), 1.0)));
    }
};

// m * x + b                              demos/line_plus_noise/line_plus_noise.py:54
struct ObeLine {
    enum { NS = 1, NP = 2, NCONS = 0, NCH = 1 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double*, double* y) {
        y[0] = obe_add(obe_mul(p[0], s[0]), p[1]);
    }
};

// Rabi counts                            demos/pipulse/pipulse.py:18-49
//   zz = ((df - fc)/B1)**2 ; f = hypot(df - fc, B1)
//   baseline*(1 - exp(-t/T1)*contrast/2*(1 - cos(pi*2*f*t))/(zz + 1))
struct ObeRabi {
    enum { NS = 2, NP = 2, NCONS = 3, NCH = 1 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double* c, double* y) {
        const double t = s[0], df = s[1], b1 = p[0], fc = p[1];
        const double det = obe_sub(df, fc);
        const double r = obe_div(det, b1);
        const double zz = obe_mul(r, r);
        const double f = hypot(det, b1);
        // numpy evaluates left to right: ((exp(-t/T1)*contrast)/2) * (1 - cos(((pi*2)*f)*t)) / (zz+1)
        const double e = obe_div(obe_mul(exp(obe_div(-t, c[2])), c[1]), 2.0);
        const double ang = obe_mul(obe_mul(6.283185307179586, f), t);
        const double osc = obe_sub(1.0, cos(ang));
        const double frac = obe_div(obe_mul(e, osc), obe_add(zz, 1.0));
        y[0] = obe_mul(c[0], obe_sub(1.0, frac));
    }
};

// (Re Z, Im Z), Z = 1/(1/(R + i w L) + i w C)   demos/lockin/lockin_of_coil.py:63-102
// Complex reciprocal as numpy's scalar/array complex division does it (Smith's method).
__device__ __forceinline__ void obe_crecip(double br, double bi, double* outr, double* outi) {
    // 1 / (br + i bi)
    const double abr = fabs(br), abi = fabs(bi);
    if (abr >= abi) {
        if (abr == 0.0 && abi == 0.0) { *outr = 1.0 / abr; *outi = 0.0 / abi; return; }
        const double rat = obe_div(bi, br);
        const double scl = obe_div(1.0, obe_add(br, obe_mul(bi, rat)));
        *outr = obe_mul(obe_add(1.0, obe_mul(0.0, rat)), scl);
        *outi = obe_mul(obe_sub(0.0, obe_mul(1.0, rat)), scl);
    } else {
        const double rat = obe_div(br, bi);
        const double scl = obe_div(1.0, obe_add(bi, obe_mul(br, rat)));
        *outr = obe_mul(obe_add(obe_mul(1.0, rat), 0.0), scl);
        *outi = obe_mul(obe_sub(obe_mul(0.0, rat), 1.0), scl);
    }
}

struct ObeLockinCoil {
    enum { NS = 1, NP = 3, NCONS = 0, NCH = 2 };
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double*, double* y) {
        const double w = s[0], L = p[0], R = p[1], C = p[2];
        // 1j*w*L : (0+1j)*w = (0*w, 1*w) then *L
        const double wl = obe_mul(w, L);
        double y1r, y1i;
        obe_crecip(R, wl, &y1r, &y1i);
        const double wc = obe_mul(w, C);
        const double tr = y1r, ti = obe_add(y1i, wc);
        obe_crecip(tr, ti, &y[0], &y[1]);
    }
};

// update-pass variant: 1/(a + ib) = (a - ib) / (a^2 + b^2) with Newton-refined reciprocals instead of the
// four IEEE divisions and two branches of numpy's complex division (the utility pass keeps the exact
// functor above, where bit-equality with the reference decides the argmax).  2-3 ulp on y.
template <>
struct ObeUpdateEval<ObeLockinCoil> {
    __device__ static __forceinline__ void eval(const double* s, const double* p, const double*, double* y) {
        const double w = s[0], L = p[0], R = p[1], C = p[2];
        const double wl = w * L;
        const double i1 = obe_rcp_fast(fma(R, R, wl * wl));
        const double tr = R * i1;                       // Re 1/(R + i wL)
        const double ti = fma(w, C, -wl * i1);          // Im 1/(R + i wL) + wC
        const double i2 = obe_rcp_fast(fma(tr, tr, ti * ti));
        y[0] = tr * i2;
        y[1] = -ti * i2;
    }
};

#endif  // OBE_MODELS_CUH
