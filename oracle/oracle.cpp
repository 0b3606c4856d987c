// oracle/oracle.cpp — TEST INFRASTRUCTURE (the parity checker), not product code.
//
// CPU restatement of the reference `dnaadjust` simultaneous solve path.  Each
// function names the reference lines it follows (tags as in SURVEY.md:
// ADJ = dynadjust/dynadjust/dnaadjust/dnaadjust.cpp, GEO = include/functions/
// dnatemplategeodesyfuncs.hpp, MFN = include/functions/dnatemplatematrixfuncs.hpp,
// MATC = include/math/dnamatrix_contiguous.cpp).
//
// The normal matrix is held packed-lower column-major exactly like matrix_2d
// (MATH:363-369) and inverted by the reference's own compiled matrix_2d when
// oracle/_ref/libref_matrix.so is loaded; otherwise by the plain Cholesky
// below ("port" mode).
#include "oracle.h"

#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

const double PI = 3.1415926535897932384626433832795029;  // dnaconsts.hpp:61
const double TWO_PI = PI + PI;
const double PRECISION_1E5 = 1.0e-5;
const double PRECISION_1E12 = 1.0e-12;
const double UNRELIABLE = 999.99;     // dnaconsts.hpp:119
const double STABLE_LIMIT = 700.;     // dnaconsts.hpp:120

std::string g_err;

// ---- optional reference-compiled helper -----------------------------------
struct RefLib {
    void* h = nullptr;
    int (*inv_packed)(double*, uint32_t) = nullptr;
    int (*inv_full)(double*, uint32_t, int) = nullptr;
    int (*mul_sym_packed)(const double*, uint32_t, const double*, double*) = nullptr;
    int (*scale_packed)(double*, uint32_t, const double*) = nullptr;
    void (*set_threads)(int) = nullptr;
} g_ref;

inline size_t pidx(uint32_t n, uint32_t i, uint32_t j)  // i >= j   (MATH:363-369)
{
    return (size_t)j * n - (size_t)j * (j - 1) / 2 + (i - j);
}
inline size_t psize(uint32_t n) { return (size_t)n * (n + 1) / 2; }

struct Ellipsoid {  // parameters/dnaellipsoid.cpp:125-135
    double a, invf, b, e2;
    Ellipsoid(double A, double INVF) : a(A), invf(INVF)
    {
        b = a * (1.0 - (1.0 / invf));
        double a2 = a * a, b2 = b * b;
        e2 = (a2 - b2) / a2;
    }
};

// parameters/dnadatumprojectionparam.hpp:63-67
inline double prime_vertical(const Ellipsoid& e, double lat)
{
    return e.a / std::sqrt(1.0 - e.e2 * (std::sin(lat) * std::sin(lat)));
}

// GEO:78-90
void geo_to_cart(const Ellipsoid& e, double lat, double lon, double h, double* X, double* Y, double* Z)
{
    double nu = prime_vertical(e, lat);
    *X = (nu + h) * std::cos(lat) * std::cos(lon);
    *Y = (nu + h) * std::cos(lat) * std::sin(lon);
    *Z = ((nu * (1. - e.e2)) + h) * std::sin(lat);
}

// GEO:154-225 (Lin & Wang Newton iteration)
void cart_to_geo(const Ellipsoid& e, double x, double y, double z, double* lat, double* lon, double* h)
{
    double p2 = (x * x) + (y * y);
    double p = std::sqrt(p2);
    double a2 = e.a * e.a;
    double b2 = e.b * e.b;
    double Z2 = z * z;
    double a2Z2 = a2 * Z2;
    double b2p2 = b2 * p2;
    double A = a2Z2 + b2p2;
    double m0 = (e.a * e.b * std::sqrt(A) * A - a2 * b2 * A) / (2. * ((a2 * a2Z2) + (b2 * b2p2)));
    double twom, a2twom, b2twom, f, df, m = m0;
    for (int i = 0; i < 5; ++i) {
        m = m0;
        twom = m * 2.;
        a2twom = a2 + twom;
        b2twom = b2 + twom;
        f = (a2 * p2 / (a2twom * a2twom)) + (b2 * Z2 / (b2twom * b2twom)) - 1.;
        if (std::fabs(f) < PRECISION_1E12)
            break;
        df = -4. * ((a2 * p2 / (a2twom * a2twom * a2twom)) + (b2 * Z2 / (b2twom * b2twom * b2twom)));
        m0 = m - (f / df);
        m = m0;
    }
    twom = m * 2.;
    double p_E = a2 * p / (a2 + twom);
    double Z_E = b2 * z / (b2 + twom);
    *lat = std::atan(a2 * Z_E / (b2 * p_E));
    *lon = std::atan(y / x);
    if (x < 0.0 && y > 0.0)
        *lon += PI;
    else if (x < 0.0 && y < 0.0)
        *lon = -(PI - *lon);
    *h = std::sqrt(((p - p_E) * (p - p_E)) + ((z - Z_E) * (z - Z_E)));
    if ((p + std::fabs(z)) < (p_E + std::fabs(Z_E)))
        *h *= -1.;
}

// ---- tiny dense helpers (3x3, column-major m[c*3+r]) ------------------------
struct M3 {
    double v[9];
    M3() { std::memset(v, 0, sizeof(v)); }
    double& operator()(int r, int c) { return v[c * 3 + r]; }
    double operator()(int r, int c) const { return v[c * 3 + r]; }
};

M3 mul(const M3& A, bool tA, const M3& B, bool tB)
{
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.;
            for (int k = 0; k < 3; ++k)
                s += (tA ? A(k, i) : A(i, k)) * (tB ? B(j, k) : B(k, j));
            C(i, j) = s;
        }
    return C;
}

// Plain lower Cholesky inverse (factor, invert the factor, W^T W) of a dense
// column-major n x n SPD matrix; full symmetric result.  "port" mode stand-in
// for dpotrf('L') + dpotri('L').
int spd_inverse_dense(double* a, uint32_t n)
{
    std::vector<double> w((size_t)n * n);
    std::memcpy(w.data(), a, (size_t)n * n * sizeof(double));
    // factor
    for (uint32_t j = 0; j < n; ++j) {
        double* cj = w.data() + (size_t)j * n;
        double d = cj[j];
        if (!(d > 0.0) || std::isnan(d))
            return 1;
        d = std::sqrt(d);
        cj[j] = d;
        double inv = 1.0 / d;
        for (uint32_t i = j + 1; i < n; ++i)
            cj[i] *= inv;
        // right-looking update of the trailing lower triangle
        for (uint32_t k = j + 1; k < n; ++k) {
            double lkj = cj[k];
            if (lkj == 0.0)
                continue;
            double* ck = w.data() + (size_t)k * n;
            for (uint32_t i = k; i < n; ++i)
                ck[i] -= cj[i] * lkj;
        }
    }
    // W = L^-1 (lower) into `inv`
    std::vector<double> iv((size_t)n * n, 0.0);
    for (uint32_t j = 0; j < n; ++j) {
        double* x = iv.data() + (size_t)j * n;
        x[j] = 1.0 / w[(size_t)j * n + j];
        for (uint32_t i = j + 1; i < n; ++i) {
            double s = 0.;
            for (uint32_t k = j; k < i; ++k)
                s += w[(size_t)k * n + i] * x[k];
            x[i] = -s / w[(size_t)i * n + i];
        }
    }
    // A^-1 = W^T W
    for (uint32_t j = 0; j < n; ++j)
        for (uint32_t i = j; i < n; ++i) {
            double s = 0.;
            const double* ci = iv.data() + (size_t)i * n;
            const double* cj = iv.data() + (size_t)j * n;
            for (uint32_t k = i; k < n; ++k)
                s += ci[k] * cj[k];
            a[(size_t)j * n + i] = s;
            a[(size_t)i * n + j] = s;
        }
    return 0;
}

// FormInverseVarianceMatrix (ADJ:8472-8517) on a small dense matrix whose upper
// (lower_is_cleared) or lower triangle is valid; result full symmetric.
int inverse_variance(double* a, uint32_t n, bool lower_is_cleared, bool use_ref)
{
    if (n == 1) {
        a[0] = 1. / a[0];
        return 0;
    }
    if (use_ref && g_ref.inv_full)
        return g_ref.inv_full(a, n, lower_is_cleared ? 1 : 0);
    if (lower_is_cleared)
        for (uint32_t j = 0; j < n; ++j)
            for (uint32_t i = j + 1; i < n; ++i)
                a[(size_t)j * n + i] = a[(size_t)i * n + j];
    return spd_inverse_dense(a, n);
}

// packed-lower inverse in place: reference path MATC:952-989, or port
int inverse_packed(std::vector<double>& ap, uint32_t n, bool use_ref)
{
    if (use_ref && g_ref.inv_packed)
        return g_ref.inv_packed(ap.data(), n);
    std::vector<double> full((size_t)n * n, 0.0);
    for (uint32_t j = 0; j < n; ++j)
        for (uint32_t i = j; i < n; ++i)
            full[(size_t)j * n + i] = ap[pidx(n, i, j)];
    int rc = spd_inverse_dense(full.data(), n);
    if (rc)
        return rc;
    for (uint32_t j = 0; j < n; ++j)
        for (uint32_t i = j; i < n; ++i)
            ap[pidx(n, i, j)] = full[(size_t)j * n + i];
    return 0;
}

// y = A x, A packed lower symmetric (dspmv, MATC:1489-1497)
void sym_packed_mv(const std::vector<double>& ap, uint32_t n, const double* x, double* y, bool use_ref)
{
    if (use_ref && g_ref.mul_sym_packed) {
        g_ref.mul_sym_packed(ap.data(), n, x, y);
        return;
    }
    for (uint32_t i = 0; i < n; ++i)
        y[i] = 0.;
    for (uint32_t j = 0; j < n; ++j) {
        const double* col = ap.data() + pidx(n, j, j);
        double xj = x[j];
        y[j] += col[0] * xj;
        double acc = 0.;
        for (uint32_t i = j + 1; i < n; ++i) {
            y[i] += col[i - j] * xj;
            acc += col[i - j] * x[i];
        }
        y[j] += acc;
    }
}

// MFN:442-479, LOCAL_TO_CART=true
M3 local_to_cart_rotation(double lat, double lon)
{
    M3 R;
    double coslat = std::cos(lat), sinlat = std::sin(lat);
    double coslon = std::cos(lon), sinlon = std::sin(lon);
    R(0, 0) = -sinlon;
    R(0, 1) = -sinlat * coslon;
    R(0, 2) = coslat * coslon;
    R(1, 0) = coslon;
    R(1, 1) = -sinlat * sinlon;
    R(1, 2) = coslat * sinlon;
    R(2, 0) = 0.;
    R(2, 1) = coslat;
    R(2, 2) = sinlat;
    return R;
}

// MFN:204-232 (geographic -> cartesian Jacobian)
M3 cart_geo_rotation(const Ellipsoid& e, double lat, double lon, double h)
{
    M3 R;
    double coslat = std::cos(lat), sinlat = std::sin(lat);
    double coslon = std::cos(lon), sinlon = std::sin(lon);
    double term1_a = e.a * e.e2;
    double one_minus_esq = 1. - e.e2;
    double nu = prime_vertical(e, lat);
    double nu_plus_h = nu + h;
    double nu_1minuse2_plus_h = nu * one_minus_esq + h;
    double term1_b = term1_a * sinlat * coslat;
    double term1_c = std::pow((1. - e.e2 * sinlat * sinlat), 1.5);
    R(0, 0) = (term1_b * coslat * coslon / term1_c) - (nu_plus_h * sinlat * coslon);
    R(0, 1) = -nu_plus_h * coslat * sinlon;
    R(0, 2) = coslat * coslon;
    R(1, 0) = (term1_b * coslat * sinlon / term1_c) - (nu_plus_h * sinlat * sinlon);
    R(1, 1) = nu_plus_h * coslat * coslon;
    R(1, 2) = coslat * sinlon;
    R(2, 0) = (term1_b * one_minus_esq * sinlat / term1_c) + (nu_1minuse2_plus_h * coslat);
    R(2, 1) = 0.;
    R(2, 2) = sinlat;
    return R;
}

// general 3x3 inverse (the reference uses sweepinverse on the Jacobian, MFN:300-313)
M3 inverse3(const M3& A)
{
    M3 B;
    double det = A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
                 A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
    double id = 1.0 / det;
    B(0, 0) = (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) * id;
    B(0, 1) = (A(0, 2) * A(2, 1) - A(0, 1) * A(2, 2)) * id;
    B(0, 2) = (A(0, 1) * A(1, 2) - A(0, 2) * A(1, 1)) * id;
    B(1, 0) = (A(1, 2) * A(2, 0) - A(1, 0) * A(2, 2)) * id;
    B(1, 1) = (A(0, 0) * A(2, 2) - A(0, 2) * A(2, 0)) * id;
    B(1, 2) = (A(0, 2) * A(1, 0) - A(0, 0) * A(1, 2)) * id;
    B(2, 0) = (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0)) * id;
    B(2, 1) = (A(0, 1) * A(2, 0) - A(0, 0) * A(2, 1)) * id;
    B(2, 2) = (A(0, 0) * A(1, 1) - A(0, 1) * A(1, 0)) * id;
    return B;
}

// ScaleGPSVCV (MFN:372-399): cart -> geographic, scale by sqrt(p,l,h), back to cart
M3 scale_gps_vcv(const Ellipsoid& e, const M3& V, double lat, double lon, double h, double pS, double lS, double hS)
{
    M3 R = cart_geo_rotation(e, lat, lon, h);
    M3 Ri = inverse3(R);
    M3 Vg = mul(mul(Ri, false, V, false), false, Ri, true);  // R^-1 V R^-T
    double s[3] = {std::sqrt(pS), std::sqrt(lS), std::sqrt(hS)};
    M3 Vs;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Vs(i, j) = s[i] * Vg(i, j) * s[j];  // ScaleMatrix: S V S^T
    return mul(mul(R, false, Vs, false), false, R, true);
}

struct Ctx {
    const oracle_opts* o;
    Ellipsoid ell;
    dna_stn_t* stn;
    uint32_t nstn;
    dna_msr_t* msr;
    uint64_t nmsr;
    uint32_t n;                    // unknowns = 3 * nstn
    std::vector<uint64_t> cml;     // first record index of every non-ignored measurement
    std::vector<double> est;       // estimated stations (3S)
    std::vector<double> ell_rows;  // measured - computed, one per design row
    std::vector<double> vinv;      // per GNSS baseline: 3x3 V^-1 (column-major), in CML order
    std::vector<double> apart;     // per scalar row: partials wrt station 1 (3) and station 2 (3)
    std::vector<double> N;         // packed normals / after Solve: packed inverse
    std::vector<double> corr;      // corrections
    std::vector<double> w;         // At V^-1 l
    uint32_t rows = 0;
    bool non_gps = false;
    bool use_ref = false;
    Ctx(const oracle_opts* O) : o(O), ell(O->semi_major, O->inv_flattening) {}
};

// LoadVarianceScaling (ADJ:4453-4491)
void load_variance_scaling(const Ctx& c, const dna_msr_t& m, double& vS, double& pS, double& lS, double& hS,
                           bool& scaleMatrix, bool& scalePartial)
{
    double lim = std::fmin(PRECISION_1E5, c.o->fixed_std_dev);
    vS = m.scale4;
    if (vS < lim)
        vS = 1.0;
    scaleMatrix = (std::fabs(vS - 1.0) > PRECISION_1E5);
    pS = m.scale1;
    lS = m.scale2;
    hS = m.scale3;
    if (pS < lim)
        pS = 1.0;
    if (lS < lim)
        lS = 1.0;
    if (hS < lim)
        hS = 1.0;
    scalePartial =
        (std::fabs(pS - 1.0) > PRECISION_1E5 || std::fabs(lS - 1.0) > PRECISION_1E5 || std::fabs(hS - 1.0) > PRECISION_1E5);
    if (scalePartial && scaleMatrix) {
        pS *= vS;
        lS *= vS;
        hS *= vS;
    }
}

// LoadVarianceMatrix_G (ADJ:4214-4309): returns V^-1 (full symmetric); first run
// writes the scaled variances back into the three records (ADJ:4281).
int load_variance_matrix_G(Ctx& c, dna_msr_t* m, M3& Vinv)
{
    double vS, pS, lS, hS;
    bool scaleMatrix, scalePartial;
    load_variance_scaling(c, m[0], vS, pS, lS, hS, scaleMatrix, scalePartial);
    M3 V;
    V(0, 0) = scaleMatrix ? m[0].term2 * vS : m[0].term2;
    V(0, 1) = scaleMatrix ? m[1].term2 * vS : m[1].term2;
    V(1, 1) = scaleMatrix ? m[1].term3 * vS : m[1].term3;
    V(0, 2) = scaleMatrix ? m[2].term2 * vS : m[2].term2;
    V(1, 2) = scaleMatrix ? m[2].term3 * vS : m[2].term3;
    V(2, 2) = scaleMatrix ? m[2].term4 * vS : m[2].term4;
    bool lowerIsClear = true;
    if (scaleMatrix || scalePartial) {
        V(1, 0) = V(0, 1);
        V(2, 0) = V(0, 2);
        V(2, 1) = V(1, 2);
        lowerIsClear = false;
    }
    if (scalePartial) {
        const dna_stn_t& s1 = c.stn[m[2].station1];
        V = scale_gps_vcv(c.ell, V, s1.currentLatitude, s1.currentLongitude, s1.currentHeight, pS, lS, hS);
    }
    if (scaleMatrix || scalePartial) {
        // SetGPSVarianceMatrix: upper triangle back into the records
        m[0].term2 = V(0, 0);
        m[1].term2 = V(0, 1);
        m[1].term3 = V(1, 1);
        m[2].term2 = V(0, 2);
        m[2].term3 = V(1, 2);
        m[2].term4 = V(2, 2);
    }
    Vinv = V;
    return inverse_variance(Vinv.v, 3, lowerIsClear, c.use_ref);
}

// FormInverseGPSVarianceMatrix (ADJ:8520-8527) for a single baseline: records as they stand
int inverse_gps_variance_G(const Ctx& c, const dna_msr_t* m, M3& Vinv)
{
    M3 V;
    V(0, 0) = m[0].term2;
    V(0, 1) = m[1].term2;
    V(1, 1) = m[1].term3;
    V(0, 2) = m[2].term2;
    V(1, 2) = m[2].term3;
    V(2, 2) = m[2].term4;
    Vinv = V;
    return inverse_variance(Vinv.v, 3, true, c.use_ref);
}

inline void lower_add(Ctx& c, uint32_t r, uint32_t col, double v)  // MATH:405-418 (packed: row >= col only)
{
    if (r < col)
        return;
    c.N[pidx(c.n, r, col)] += v;
}

// scan the record list into the CML (first record of each non-ignored measurement)
int build_cml(Ctx& c)
{
    uint64_t i = 0;
    while (i < c.nmsr) {
        const dna_msr_t& m = c.msr[i];
        uint64_t step = 1;
        switch (m.measType) {
        case 'G':
            step = 3;
            break;
        case 'X':
        case 'Y': {
            // cluster: vectorCount1 members, each 3 records + 3 * vectorCount2(member) covariance records
            uint64_t j = i;
            uint32_t members = m.vectorCount1;
            for (uint32_t k = 0; k < members; ++k)
                j += 3 + 3ull * c.msr[j].vectorCount2;
            step = j - i;
            break;
        }
        case 'D':
            step = 1ull + m.vectorCount1;
            break;
        default:
            step = 1;
        }
        if (!m.ignore) {
            if (m.measType != 'G' && m.measType != 'S' && m.measType != 'L') {
                g_err = std::string("oracle: measurement type '") + m.measType + "' not restated yet";
                return 3;
            }
            if (m.measType != 'G')
                c.non_gps = true;
            c.cml.push_back(i);
        }
        i += step;
    }
    return 0;
}

// EllipsoidHeight (GEO:909-920)
double ellipsoid_height(const Ellipsoid& e, double X, double Y, double Z, double lat, double* nu, double* Zn)
{
    *nu = prime_vertical(e, lat);
    *Zn = e.e2 * (*nu) * std::sin(lat);
    return std::sqrt(X * X + Y * Y + std::pow(Z + (*Zn), 2)) - (*nu);
}

// FillDesignNormalMeasurementsMatrices (ADJ:3888-4055) for the types restated.
//   build=true : first pass — l, partials, V^-1 (with variance scaling write-back), first-run reductions
//   build=false: re-linearise — l (and the partials of the non-GNSS rows; GNSS design never changes, ADJ:5294-5301)
int fill_design_normals(Ctx& c, bool build)
{
    uint32_t row = 0;
    size_t g = 0, sc = 0;
    for (uint64_t first : c.cml) {
        dna_msr_t* m = &c.msr[first];
        switch (m->measType) {
        case 'G': {
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            // UpdateDesignMeasMatrices_GX (ADJ:5283-5350)
            for (int r = 0; r < 3; ++r) {
                c.ell_rows[row + r] = m[r].term1 - (c.est[s2 + r] - c.est[s1 + r]);
                if (build)
                    m[r].preAdjMeas = m[r].term1;
            }
            if (build) {
                M3 Vinv;
                int rc = load_variance_matrix_G(c, m, Vinv);
                if (rc) {
                    g_err = "oracle: GNSS variance matrix inversion failed";
                    return rc;
                }
                std::memcpy(&c.vinv[g * 9], Vinv.v, sizeof(Vinv.v));
            }
            row += 3;
            ++g;
            break;
        }
        case 'S': {
            // UpdateDesignNormalMeasMatrices_S (ADJ:5437-5493)
            if (build)
                m->preAdjMeas = m->term1;
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            const dna_stn_t& st1 = c.stn[m->station1];
            double cl = std::cos(st1.currentLatitude), sl = std::sin(st1.currentLatitude);
            double co = std::cos(st1.currentLongitude), so = std::sin(st1.currentLongitude);
            // CartesianElementsFromInstrumentHeight (GEO:763-771): both heights are rotated at station 1
            double dXih = cl * co * m->term3, dYih = cl * so * m->term3, dZih = sl * m->term3;
            double dXth = cl * co * m->term4, dYth = cl * so * m->term4, dZth = sl * m->term4;
            double dX = c.est[s2] - c.est[s1] + dXth - dXih;
            double dY = c.est[s2 + 1] - c.est[s1 + 1] + dYth - dYih;
            double dZ = c.est[s2 + 2] - c.est[s1 + 2] + dZth - dZih;
            double comp = std::sqrt(dX * dX + dY * dY + dZ * dZ);
            c.ell_rows[row] = m->term1 - comp;
            double* a = &c.apart[sc * 6];
            a[0] = -dX / comp;
            a[1] = -dY / comp;
            a[2] = -dZ / comp;
            a[3] = -a[0];
            a[4] = -a[1];
            a[5] = -a[2];
            row += 1;
            ++sc;
            break;
        }
        case 'L': {
            // UpdateDesignNormalMeasMatrices_L (ADJ:5717-5784)
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            const dna_stn_t& st1 = c.stn[m->station1];
            const dna_stn_t& st2 = c.stn[m->station2];
            double nu1, nu2, Zn1, Zn2;
            double h1 = ellipsoid_height(c.ell, c.est[s1], c.est[s1 + 1], c.est[s1 + 2], st1.currentLatitude, &nu1, &Zn1);
            double h2 = ellipsoid_height(c.ell, c.est[s2], c.est[s2 + 1], c.est[s2 + 2], st2.currentLatitude, &nu2, &Zn2);
            double comp = h2 - h1;
            if (build) {
                m->preAdjMeas = m->term1;   // InitialiseMeasurement (ADJ:3913-3935)
                if (std::fabs(st1.geoidSep) > 1.0e-4 || std::fabs(st2.geoidSep) > 1.0e-4) {
                    m->preAdjCorr = st2.geoidSep - st1.geoidSep;
                    m->term1 += m->preAdjCorr;
                }
            }
            c.ell_rows[row] = m->term1 - comp;
            double* a = &c.apart[sc * 6];
            a[0] = -c.est[s1] / (nu1 + h1);
            a[1] = -c.est[s1 + 1] / (nu1 + h1);
            a[2] = -(c.est[s1 + 2] + Zn1) / (nu1 + h1);
            a[3] = c.est[s2] / (nu2 + h2);
            a[4] = c.est[s2 + 1] / (nu2 + h2);
            a[5] = (c.est[s2 + 2] + Zn2) / (nu2 + h2);
            row += 1;
            ++sc;
            break;
        }
        default:
            break;
        }
    }
    return 0;
}

// UpdateNormals (ADJ:1364-1455): N from the stored At V^-1 / design of every measurement
void update_normals(Ctx& c)
{
    size_t g = 0, sc = 0;
    for (uint64_t first : c.cml) {
        const dna_msr_t* m = &c.msr[first];
        uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
        if (m->measType == 'G') {
            // UpdateNormals_G (ADJ:1664-1684) through add_normal_3x3_from_atvinv_columns (ADJ:1478-1491):
            // AtVinv[s1.., rows] = -V^-1 ; AtVinv[s2.., rows] = +V^-1
            M3 Vinv;
            std::memcpy(Vinv.v, &c.vinv[g * 9], sizeof(Vinv.v));
            for (int col = 0; col < 3; ++col)
                for (int r = 0; r < 3; ++r) {
                    lower_add(c, s2 + r, s2 + col, 1. * Vinv(r, col));
                    lower_add(c, s1 + r, s1 + col, -1. * (-Vinv(r, col)));
                    lower_add(c, s1 + r, s2 + col, 1. * (-Vinv(r, col)));
                    lower_add(c, s2 + r, s1 + col, -1. * Vinv(r, col));
                }
            ++g;
        } else {
            // UpdateNormals_BCEKLMSVZ (ADJ:1640-1651): At V^-1 = a / term2 (UpdateAtVinv, ADJ:1285-1320)
            const double* a = &c.apart[sc * 6];
            const double p = 1. / m->term2;
            const uint32_t st[2] = {s1, s2};
            for (int bi = 0; bi < 2; ++bi)
                for (int bj = 0; bj < 2; ++bj)
                    for (int col = 0; col < 3; ++col)
                        for (int r = 0; r < 3; ++r)
                            lower_add(c, st[bi] + r, st[bj] + col, (p * a[3 * bi + r]) * a[3 * bj + col]);
            ++sc;
        }
    }
}

// FormConstraintStationVarianceMatrix (ADJ:2041-2137) -> inverse variance block
int constraint_block(const Ctx& c, const dna_stn_t& s, M3& out)
{
    double varC = c.o->fixed_std_dev * c.o->fixed_std_dev;
    double varF = c.o->free_std_dev * c.o->free_std_dev;
    const char* k = s.stationConst;
    out = M3();
    if (k[0] == 'C' && k[1] == 'C' && k[2] == 'C') {
        out(0, 0) = out(1, 1) = out(2, 2) = 1. / varC;
        return 0;
    }
    if (k[0] == 'F' && k[1] == 'F' && k[2] == 'F') {
        out(0, 0) = out(1, 1) = out(2, 2) = 1. / varF;
        return 0;
    }
    M3 L;
    bool llh = (s.suppliedStationType == DNA_LLH_TYPE || s.suppliedStationType == DNA_LLh_TYPE);
    double v0 = (k[0] == 'F') ? varF : varC;
    double v1 = (k[1] == 'F') ? varF : varC;
    if (llh) {
        L(1, 1) = v0;  // latitude -> north
        L(0, 0) = v1;  // longitude -> east
    } else {
        L(0, 0) = v0;
        L(1, 1) = v1;
    }
    L(2, 2) = (k[2] == 'F') ? varF : varC;
    M3 V;
    if (s.suppliedStationType == DNA_XYZ_TYPE)
        V = L;
    else {
        M3 R = local_to_cart_rotation(s.currentLatitude, s.currentLongitude);
        V = mul(mul(R, false, L, false), false, R, true);  // MFN:592-621
    }
    out = V;
    return inverse_variance(out.v, 3, false, c.use_ref);
}

// AddConstraintStationstoNormalsSimultaneous (ADJ:2010-2037)
int add_constraints(Ctx& c)
{
    for (uint32_t s = 0; s < c.nstn; ++s) {
        M3 B;
        int rc = constraint_block(c, c.stn[s], B);
        if (rc)
            return rc;
        for (int col = 0; col < 3; ++col)
            for (int r = col; r < 3; ++r)  // blockadd into packed keeps row >= col (MATC:1161-1172)
                c.N[pidx(c.n, s * 3 + r, s * 3 + col)] += B(r, col);
    }
    return 0;
}

// At V^-1 l (the reference forms dense AtVinv and calls dgemm, ADJ:6659-6660)
void weighted_rhs(Ctx& c)
{
    std::fill(c.w.begin(), c.w.end(), 0.0);
    uint32_t row = 0;
    size_t g = 0, sc = 0;
    for (uint64_t first : c.cml) {
        const dna_msr_t* m = &c.msr[first];
        if (m->measType == 'G') {
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            const double* V = &c.vinv[g * 9];
            for (int r = 0; r < 3; ++r) {
                double t = 0.;
                for (int k = 0; k < 3; ++k)
                    t += V[k * 3 + r] * c.ell_rows[row + k];
                c.w[s1 + r] += -t;
                c.w[s2 + r] += t;
            }
            row += 3;
            ++g;
        } else {
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            const double* a = &c.apart[sc * 6];
            const double p = 1. / m->term2;
            for (int r = 0; r < 3; ++r) {
                c.w[s1 + r] += (p * a[r]) * c.ell_rows[row];
                c.w[s2 + r] += (p * a[3 + r]) * c.ell_rows[row];
            }
            row += 1;
            ++sc;
        }
    }
}

// Solve (ADJ:6586-6667)
int solve(Ctx& c, bool compute_inverse, double* t_inv)
{
    if (compute_inverse) {
        std::vector<double> sdiag;
        if (c.o->scale_normals_to_unity) {
            sdiag.resize(c.n);
            for (uint32_t i = 0; i < c.n; ++i)
                sdiag[i] = 1.0 / std::sqrt(c.N[pidx(c.n, i, i)]);
            for (uint32_t j = 0; j < c.n; ++j)
                for (uint32_t i = j; i < c.n; ++i)
                    c.N[pidx(c.n, i, j)] *= sdiag[i] * sdiag[j];
        }
        auto t0 = std::chrono::steady_clock::now();
        int rc = inverse_packed(c.N, c.n, c.use_ref);
        *t_inv += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc) {
            g_err = "Matrix inversion failed, the matrix is singular.";  // MATC:983
            return rc;
        }
        if (std::isnan(c.N[0]) || std::isinf(c.N[0])) {
            g_err = "Solve(): Invalid variance matrix";
            return 4;
        }
        if (c.o->scale_normals_to_unity)
            for (uint32_t j = 0; j < c.n; ++j)
                for (uint32_t i = j; i < c.n; ++i)
                    c.N[pidx(c.n, i, j)] *= sdiag[i] * sdiag[j];
    }
    weighted_rhs(c);
    sym_packed_mv(c.N, c.n, c.w.data(), c.corr.data(), c.use_ref);
    return 0;
}

// Precision_Adjusted_GNSS_bsl (MFN:255-297): upper triangle of A Q A^T for A = [-I  I]
void precision_adjusted_gnss_bsl(const Ctx& c, uint32_t s1, uint32_t s2, double out6[6])
{
    auto Q = [&](uint32_t i, uint32_t j) { return i >= j ? c.N[pidx(c.n, i, j)] : c.N[pidx(c.n, j, i)]; };
    double tmp[3][6];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double t = 0.;
            t += -Q(s1 + i, s1 + j);
            t += Q(s2 + i, s1 + j);
            tmp[i][j] = t;
        }
        for (int j = 0; j < 3; ++j) {
            double t = 0.;
            t += -Q(s1 + i, s2 + j);
            t += Q(s2 + i, s2 + j);
            tmp[i][3 + j] = t;
        }
    }
    int k = 0;
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            out6[k++] = tmp[i][j + 3] - tmp[i][j];
}

// UpdateMsrRecord + UpdateMsrRecordStats (ADJ:8187-8298)
void update_msr_record(dna_msr_t& m, double mmc, double adjPrec, double measPrec, double critical, uint32_t& outliers)
{
    m.measCorr = -mmc;
    m.measAdj = m.term1 + m.measCorr;
    m.measAdjPrec = adjPrec;
    m.residualPrec = measPrec - m.measAdjPrec;
    if (m.residualPrec < 0.0)
        m.residualPrec = std::fabs(m.residualPrec);
    m.PelzerRel = std::sqrt(measPrec) / std::sqrt(m.residualPrec);
    if (m.PelzerRel < 0. || m.PelzerRel > STABLE_LIMIT)
        m.PelzerRel = UNRELIABLE;
    m.NStat = m.measCorr / std::sqrt(m.residualPrec);
    if (std::fabs(m.NStat) > critical)
        outliers++;
}

// inverse standard-normal CDF (the reference uses boost::math::quantile, ADJ:203-206);
// Acklam's rational approximation refined by one Halley step on erfc — ~1e-15.
double norm_quantile(double p)
{
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02,  -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01,  -1.328068155288572e+01};
    static const double cc[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00,  2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double q, r, x;
    if (p < 0.02425) {
        q = std::sqrt(-2 * std::log(p));
        x = (((((cc[0] * q + cc[1]) * q + cc[2]) * q + cc[3]) * q + cc[4]) * q + cc[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    } else if (p <= 1 - 0.02425) {
        q = p - 0.5;
        r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        q = std::sqrt(-2 * std::log(1 - p));
        x = -(((((cc[0] * q + cc[1]) * q + cc[2]) * q + cc[3]) * q + cc[4]) * q + cc[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }
    double e = 0.5 * std::erfc(-x / std::sqrt(2.0)) - p;
    double u = e * std::sqrt(2 * PI) * std::exp(x * x / 2);
    x = x - u / (1 + x * u / 2);
    return x;
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

extern "C" {

const char* oracle_last_error(void) { return g_err.c_str(); }

void oracle_default_opts(oracle_opts* o)
{
    o->fixed_std_dev = 1.0e-6;
    o->free_std_dev = 10.0;
    o->iteration_threshold = (double)0.0005f;
    o->semi_major = 6378137.0;
    o->inv_flattening = 298.257222101;
    o->confidence_interval = 95.0;
    o->max_iterations = 10;
    o->scale_normals_to_unity = 0;
    o->use_ref = 1;
    o->threads = 0;
}

int oracle_load_ref(const char* path)
{
    if (g_ref.h)
        return 1;
    void* h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        g_err = std::string("oracle_load_ref: ") + dlerror();
        return 0;
    }
    g_ref.h = h;
    g_ref.inv_packed = (int (*)(double*, uint32_t))dlsym(h, "ref_cholesky_inverse_packed");
    g_ref.inv_full = (int (*)(double*, uint32_t, int))dlsym(h, "ref_cholesky_inverse_full");
    g_ref.mul_sym_packed = (int (*)(const double*, uint32_t, const double*, double*))dlsym(h, "ref_multiply_sym_packed");
    g_ref.scale_packed = (int (*)(double*, uint32_t, const double*))dlsym(h, "ref_scale_symmetric_diagonal_packed");
    g_ref.set_threads = (void (*)(int))dlsym(h, "ref_set_threads");
    if (!g_ref.inv_packed || !g_ref.inv_full || !g_ref.mul_sym_packed) {
        g_err = "oracle_load_ref: missing symbols";
        dlclose(h);
        g_ref = RefLib();
        return 0;
    }
    return 1;
}

int oracle_ref_loaded(void) { return g_ref.h != nullptr; }

void oracle_geo_to_cart(double lat, double lon, double h, double a, double invf, double* xyz)
{
    Ellipsoid e(a, invf);
    geo_to_cart(e, lat, lon, h, &xyz[0], &xyz[1], &xyz[2]);
}

void oracle_cart_to_geo(double x, double y, double z, double a, double invf, double* llh)
{
    Ellipsoid e(a, invf);
    cart_to_geo(e, x, y, z, &llh[0], &llh[1], &llh[2]);
}

int oracle_spd_inverse(double* a, uint32_t n, int use_ref)
{
    if (use_ref && g_ref.inv_packed) {
        std::vector<double> ap(psize(n));
        for (uint32_t j = 0; j < n; ++j)
            for (uint32_t i = j; i < n; ++i)
                ap[pidx(n, i, j)] = a[(size_t)j * n + i];
        int rc = g_ref.inv_packed(ap.data(), n);
        if (rc)
            return rc;
        for (uint32_t j = 0; j < n; ++j)
            for (uint32_t i = j; i < n; ++i)
                a[(size_t)j * n + i] = a[(size_t)i * n + j] = ap[pidx(n, i, j)];
        return 0;
    }
    return spd_inverse_dense(a, n);
}

int oracle_adjust_simultaneous(const oracle_opts* opts, dna_stn_t* stn, uint32_t nstn, dna_msr_t* msr, uint64_t nmsr,
                               double* est_xyz, double* normals_full, double* rhs, double* first_corr, double* vcv_full,
                               oracle_result* res)
{
    g_err.clear();
    Ctx c(opts);
    c.stn = stn;
    c.nstn = nstn;
    c.msr = msr;
    c.nmsr = nmsr;
    c.n = nstn * 3;
    c.use_ref = opts->use_ref && g_ref.h;
    if (c.use_ref && g_ref.set_threads && opts->threads > 0)
        g_ref.set_threads(opts->threads);
    std::memset(res, 0, sizeof(*res));
    res->used_ref = c.use_ref ? 1 : 0;

    double t0 = now_s();
    // InitialiseAdjustment (ADJ:198-246)
    double conf = opts->confidence_interval * 0.01;
    conf += (1.0 - conf) / 2.0;
    double critical = norm_quantile(conf);
    res->critical_value = critical;

    int rc = build_cml(c);
    if (rc)
        return rc;
    for (uint64_t first : c.cml)
        c.rows += (msr[first].measType == 'G') ? 3 : 1;

    // PopulateEstimatedStationMatrix (ADJ:632-693)
    c.est.resize(c.n);
    uint32_t unknownParams = c.n;
    for (uint32_t s = 0; s < nstn; ++s) {
        geo_to_cart(c.ell, stn[s].currentLatitude, stn[s].currentLongitude, stn[s].currentHeight, &c.est[3 * s],
                    &c.est[3 * s + 1], &c.est[3 * s + 2]);
        for (int k = 0; k < 3; ++k)
            if (stn[s].stationConst[k] == 'C')
                unknownParams--;
    }
    c.N.assign(psize(c.n), 0.0);
    c.ell_rows.assign(c.rows, 0.0);
    c.vinv.assign(c.cml.size() * 9, 0.0);
    c.apart.assign(c.cml.size() * 6, 0.0);
    c.corr.assign(c.n, 0.0);
    c.w.assign(c.n, 0.0);

    rc = fill_design_normals(c, true);
    if (rc)
        return rc;
    update_normals(c);
    rc = add_constraints(c);
    if (rc)
        return rc;
    res->seconds_prepare = now_s() - t0;

    if (normals_full)
        for (uint32_t j = 0; j < c.n; ++j)
            for (uint32_t i = j; i < c.n; ++i)
                normals_full[(size_t)j * c.n + i] = normals_full[(size_t)i * c.n + j] = c.N[pidx(c.n, i, j)];

    // AdjustSimultaneous (ADJ:2413-2511)
    double maxCorr = 0.;
    uint32_t maxRow = 0;
    uint32_t iter = 0;
    for (uint32_t i = 0; i < opts->max_iterations; ++i) {
        ++iter;
        double ts = now_s();
        rc = solve(c, iter < 2 || c.non_gps, &res->seconds_inverse);
        res->seconds_solve += now_s() - ts;
        if (rc)
            return rc;
        if (iter == 1) {
            if (rhs)
                std::memcpy(rhs, c.w.data(), c.n * sizeof(double));
            if (first_corr)
                std::memcpy(first_corr, c.corr.data(), c.n * sizeof(double));
        }
        for (uint32_t k = 0; k < c.n; ++k)
            c.est[k] += c.corr[k];
        // compute_maximum_value (MATC:1532-1555): first element of largest magnitude
        maxRow = 0;
        for (uint32_t k = 0; k < c.n; ++k)
            if (std::fabs(c.corr[k]) > std::fabs(c.corr[maxRow]))
                maxRow = k;
        maxCorr = c.corr[maxRow];
        bool iterate = std::fabs(maxCorr) > opts->iteration_threshold;
        if (!iterate)
            break;
        bool lastIteration = (i + 1 >= opts->max_iterations);
        // UpdateAdjustment(!lastIteration) (ADJ:473-627)
        if (c.non_gps || lastIteration)   // UpdateGeographicCoords (ADJ:543-545, ADJ:8734)
            for (uint32_t s = 0; s < nstn; ++s)
                cart_to_geo(c.ell, c.est[3 * s], c.est[3 * s + 1], c.est[3 * s + 2], &stn[s].currentLatitude,
                            &stn[s].currentLongitude, &stn[s].currentHeight);
        rc = fill_design_normals(c, false);
        if (rc)
            return rc;
        if (!lastIteration && c.non_gps) {
            // partials changed: rebuild the normals (ADJ:583-589)
            std::fill(c.N.begin(), c.N.end(), 0.0);
            update_normals(c);
            rc = add_constraints(c);
            if (rc)
                return rc;
        }
    }
    res->iterations = iter;
    res->max_corr = maxCorr;
    res->max_corr_row = maxRow;
    res->converged = std::fabs(maxCorr) <= opts->iteration_threshold;

    // GenerateStatistics (ADJ:6802-6841): UpdateAdjustment(false) -> UpdateGeographicCoords (ADJ:8734) + l
    for (uint32_t s = 0; s < nstn; ++s)
        cart_to_geo(c.ell, c.est[3 * s], c.est[3 * s + 1], c.est[3 * s + 2], &stn[s].currentLatitude,
                    &stn[s].currentLongitude, &stn[s].currentHeight);
    rc = fill_design_normals(c, false);
    if (rc)
        return rc;

    // ComputeStatistics (ADJ:7116-7148)
    uint32_t outliers = 0;
    double chi = 0.;
    {
        uint32_t row = 0;
        size_t sc = 0;
        auto Q = [&](uint32_t i, uint32_t j) { return i >= j ? c.N[pidx(c.n, i, j)] : c.N[pidx(c.n, j, i)]; };
        for (uint64_t first : c.cml) {
            dna_msr_t* m = &msr[first];
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            if (m->measType != 'G') {
                // ComputePrecisionAdjMsrs_BCEKLMSVZ (ADJ:7949-7982): a Q a^T over the two stations
                const double* a = &c.apart[sc * 6];
                const uint32_t st[2] = {s1, s2};
                double prec = 0.;
                for (int bs = 0; bs < 2; ++bs)
                    for (int i = 0; i < 3; ++i) {
                        double part = 0.;
                        for (int bj = 0; bj < 2; ++bj)
                            for (int k = 0; k < 3; ++k)
                                part += a[3 * bj + k] * Q(st[bj] + k, st[bs] + i);
                        prec += part * a[3 * bs + i];
                    }
                update_msr_record(*m, c.ell_rows[row], prec, m->term2, critical, outliers);
                if (m->measType == 'L')
                    m->measAdj -= m->preAdjCorr;   // ADJ:8241-8244
                chi += c.ell_rows[row] * c.ell_rows[row] / m->term2;   // ADJ:8430-8437
                row += 1;
                ++sc;
                continue;
            }
            // ComputePrecisionAdjMsrs_GX (ADJ:8006-8032)
            double p6[6];
            precision_adjusted_gnss_bsl(c, s1, s2, p6);
            // UpdateMsrRecords_GXY (ADJ:8152-8184): XX row+0, YY row+3, ZZ row+5
            update_msr_record(m[0], c.ell_rows[row + 0], p6[0], m[0].term2, critical, outliers);
            update_msr_record(m[1], c.ell_rows[row + 1], p6[3], m[1].term3, critical, outliers);
            update_msr_record(m[2], c.ell_rows[row + 2], p6[5], m[2].term4, critical, outliers);
            // ComputeChiSquare_G (ADJ:8530-8549)
            M3 Vinv;
            rc = inverse_gps_variance_G(c, m, Vinv);
            if (rc)
                return rc;
            double cs = 0.;
            for (int r = 0; r < 3; ++r)
                for (int col = 0; col < 3; ++col)
                    cs += Vinv(r, col) * c.ell_rows[row + r] * c.ell_rows[row + col];
            chi += cs;
            row += 3;
        }
    }
    res->chi_squared = chi;
    res->measurement_params = c.rows;
    res->unknown_params = unknownParams;
    res->dof = (int64_t)c.rows - (int64_t)unknownParams;  // ADJ:6856
    res->sigma_zero = res->dof != 0 ? chi / (double)res->dof : 0.;
    res->outliers = outliers;
    // ComputeGlobalPelzer (ADJ:8302-8427)
    {
        double sum = 0.;
        uint32_t num = 0;
        for (uint64_t first : c.cml) {
            dna_msr_t* m = &msr[first];
            if (m->measType != 'G') {
                if (m->PelzerRel > 0. && m->PelzerRel < STABLE_LIMIT) {   // ADJ:8339-8345
                    sum += (m->PelzerRel * m->PelzerRel - 1.);
                    num++;
                } else
                    m->PelzerRel = UNRELIABLE;
                continue;
            }
            for (int k = 0; k < 3; ++k) {
                if (m[k].PelzerRel > 0. && m[k].PelzerRel < UNRELIABLE) {
                    sum += (m[k].PelzerRel * m[k].PelzerRel - 1.);
                    num++;
                } else
                    m[k].PelzerRel = UNRELIABLE;
            }
        }
        res->global_pelzer = num > 0 ? std::sqrt(sum / num) : UNRELIABLE;
    }

    if (est_xyz)
        std::memcpy(est_xyz, c.est.data(), c.n * sizeof(double));
    if (vcv_full)
        for (uint32_t j = 0; j < c.n; ++j)
            for (uint32_t i = j; i < c.n; ++i)
                vcv_full[(size_t)j * c.n + i] = vcv_full[(size_t)i * c.n + j] = c.N[pidx(c.n, i, j)];
    return 0;
}

}  // extern "C"
