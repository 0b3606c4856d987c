// geodesy.h — closed-form geodesy shared by host code and CUDA kernels.
// Formulas follow the reference's header templates (cited per function); the
// code is written for this engine (plain structs, no matrix class).
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define GADJ_HD __host__ __device__ __forceinline__
#else
#define GADJ_HD inline
#endif

namespace gadj {

constexpr double kPi = 3.1415926535897932384626433832795029;  // dnaconsts.hpp:61

struct Ellipsoid {  // parameters/dnaellipsoid.cpp:125-135
    double a, b, e2;
};

GADJ_HD Ellipsoid make_ellipsoid(double a, double invf)
{
    Ellipsoid e;
    e.a = a;
    e.b = a * (1.0 - (1.0 / invf));
    double a2 = a * a, b2 = e.b * e.b;
    e.e2 = (a2 - b2) / a2;
    return e;
}

// prime vertical radius (dnadatumprojectionparam.hpp:63-67)
GADJ_HD double prime_vertical(const Ellipsoid& e, double lat)
{
    double s = sin(lat);
    return e.a / sqrt(1.0 - e.e2 * (s * s));
}

// GeoToCart (dnatemplategeodesyfuncs.hpp:78-90)
GADJ_HD void geo_to_cart(const Ellipsoid& e, double lat, double lon, double h, double* xyz)
{
    double nu = prime_vertical(e, lat);
    xyz[0] = (nu + h) * cos(lat) * cos(lon);
    xyz[1] = (nu + h) * cos(lat) * sin(lon);
    xyz[2] = ((nu * (1. - e.e2)) + h) * sin(lat);
}

// CartToGeo, Lin & Wang Newton iteration (dnatemplategeodesyfuncs.hpp:154-225)
GADJ_HD void cart_to_geo(const Ellipsoid& e, double x, double y, double z, double* llh)
{
    double p2 = (x * x) + (y * y);
    double p = sqrt(p2);
    double a2 = e.a * e.a, b2 = e.b * e.b;
    double Z2 = z * z;
    double a2Z2 = a2 * Z2, b2p2 = b2 * p2;
    double A = a2Z2 + b2p2;
    double m0 = (e.a * e.b * sqrt(A) * A - a2 * b2 * A) / (2. * ((a2 * a2Z2) + (b2 * b2p2)));
    double m = m0;
    for (int i = 0; i < 5; ++i) {
        m = m0;
        double twom = m * 2.;
        double a2t = a2 + twom, b2t = b2 + twom;
        double f = (a2 * p2 / (a2t * a2t)) + (b2 * Z2 / (b2t * b2t)) - 1.;
        if (fabs(f) < 1.0e-12)
            break;
        double df = -4. * ((a2 * p2 / (a2t * a2t * a2t)) + (b2 * Z2 / (b2t * b2t * b2t)));
        m0 = m - (f / df);
        m = m0;
    }
    double twom = m * 2.;
    double pE = a2 * p / (a2 + twom);
    double zE = b2 * z / (b2 + twom);
    llh[0] = atan(a2 * zE / (b2 * pE));
    double lon = atan(y / x);
    if (x < 0.0 && y > 0.0)
        lon += kPi;
    else if (x < 0.0 && y < 0.0)
        lon = -(kPi - lon);
    llh[1] = lon;
    double h = sqrt(((p - pE) * (p - pE)) + ((z - zE) * (z - zE)));
    if ((p + fabs(z)) < (pE + fabs(zE)))
        h *= -1.;
    llh[2] = h;
}

// local (e,n,up) -> Cartesian rotation, row-major R[r*3+c]  (dnatemplatematrixfuncs.hpp:442-479)
GADJ_HD void local_to_cart_rotation(double lat, double lon, double* R)
{
    double cl = cos(lat), sl = sin(lat), co = cos(lon), so = sin(lon);
    R[0] = -so;
    R[1] = -sl * co;
    R[2] = cl * co;
    R[3] = co;
    R[4] = -sl * so;
    R[5] = cl * so;
    R[6] = 0.;
    R[7] = cl;
    R[8] = sl;
}

// inverse of a symmetric positive-definite 3x3 given by its upper triangle
// (xx, xy, xz, yy, yz, zz) via Cholesky; out6 in the same order.  returns false when not SPD.
GADJ_HD bool spd3_inverse(const double* v, double* out6)
{
    // A = L L^T
    double l00 = v[0];
    if (!(l00 > 0.0))
        return false;
    l00 = sqrt(l00);
    double l10 = v[1] / l00, l20 = v[2] / l00;
    double d1 = v[3] - l10 * l10;
    if (!(d1 > 0.0))
        return false;
    double l11 = sqrt(d1);
    double l21 = (v[4] - l20 * l10) / l11;
    double d2 = v[5] - l20 * l20 - l21 * l21;
    if (!(d2 > 0.0))
        return false;
    double l22 = sqrt(d2);
    // W = L^-1 (lower)
    double w00 = 1.0 / l00, w11 = 1.0 / l11, w22 = 1.0 / l22;
    double w10 = -l10 * w00 * w11;
    double w21 = -l21 * w11 * w22;
    double w20 = -(l20 * w00 + l21 * w10) * w22;
    // A^-1 = W^T W
    out6[0] = w00 * w00 + w10 * w10 + w20 * w20;
    out6[1] = w10 * w11 + w20 * w21;
    out6[2] = w20 * w22;
    out6[3] = w11 * w11 + w21 * w21;
    out6[4] = w21 * w22;
    out6[5] = w22 * w22;
    return true;
}

}  // namespace gadj
