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
//
// Pinned by the reference's own end-to-end expected outputs (tests/test_golden.py):
// sampleData/gnss.simult.adj.expected (every printed digit) and
// sampleData/urban.phased.adj.expected (every terrestrial type; to the accuracy
// of the exported geoid values), through tests/golden/*.npz; the reference's
// matrix_2d, compiled in place, inverts the normals.
#include "oracle.h"

#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <memory>
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

#include "oracle_types.h"

// one entry of the CML (ADJH:1216-1218): a measurement or a cluster
struct Meas {
    uint64_t first = 0;
    char type = 0;
    uint32_t row0 = 0, nrows = 0;
    size_t g = 0;                  // G: index into Ctx::vinv
    std::vector<uint64_t> rec;     // D: direction record that stores each derived angle; X/Y: first record of each member
    std::vector<uint64_t> base;    // D: record that supplies station1/station2/term3/term4 of each angle (RO, then the previous direction)
    std::vector<double> vinv;      // D/X/Y: dense nrows x nrows V^-1, column-major, full
};

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
    std::vector<Meas> meas;        // parallel to cml
    std::vector<Row> row;          // per design row of the non-G measurements: stations and partials
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

// scan the record list into the CML (first record of each non-ignored measurement / cluster)
int build_cml(Ctx& c)
{
    uint64_t i = 0;
    while (i < c.nmsr) {
        const dna_msr_t& m = c.msr[i];
        uint64_t step = 1;
        Meas me;
        me.first = i;
        me.type = m.measType;
        switch (m.measType) {
        case 'G':
            step = 3;
            me.nrows = 3;
            break;
        case 'X':
        case 'Y': {
            // cluster: vectorCount1 members, each 3 records + 3 * vectorCount2(member) covariance records
            uint64_t j = i;
            uint32_t members = m.vectorCount1;
            for (uint32_t k = 0; k < members; ++k) {
                me.rec.push_back(j);
                j += 3 + 3ull * c.msr[j].vectorCount2;
            }
            step = j - i;
            me.nrows = 3 * members;
            break;
        }
        case 'D': {
            // RO record + target directions: vectorCount1 records in all (dnadirectionset.cpp:430-466);
            // vectorCount2 of them are not ignored -> vectorCount2 - 1 derived angles (ADJ:5088-5090)
            step = m.vectorCount1 ? m.vectorCount1 : 1;
            uint64_t prev = i;
            for (uint64_t j = i + 1; j < i + step && me.rec.size() + 1 < m.vectorCount2; ++j) {
                if (c.msr[j].ignore)
                    continue;   // ADJ:5120-5129
                me.rec.push_back(j);
                me.base.push_back(prev);
                prev = j;
            }
            me.nrows = (uint32_t)me.rec.size();
            break;
        }
        default:
            step = 1;
            me.nrows = 1;
        }
        if (!m.ignore && me.nrows > 0) {
            if (m.measType != 'G')
                c.non_gps = c.non_gps || (m.measType != 'X' && m.measType != 'Y');   // ContainsNonGPS (msr tally)
            c.cml.push_back(i);
            c.meas.push_back(std::move(me));
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

// LoadVarianceMatrix_D (ADJ:4059-4188): tridiagonal variance matrix of the angles derived from a round of
// directions; first run stores variance / covariance in scale2 / scale3 of the direction records
// (SetDirectionsVarianceMatrix, MFN:126-160), later runs read them back (GetDirectionsVarianceMatrix, MFN:53-82).
int load_variance_matrix_D(Ctx& c, Meas& me, bool build)
{
    const uint32_t n = me.nrows;
    std::vector<double> V((size_t)n * n, 0.0);
    if (build) {
        double previousVariance = c.msr[me.first].term2;
        for (uint32_t a = 0; a < n; ++a) {
            const double var = c.msr[me.rec[a]].term2;
            // A = [-1 1] per angle; AV = [-prev, var]
            V[(size_t)a * n + a] += (previousVariance * -1) * -1;
            V[(size_t)a * n + a] += var * 1;
            if (a + 1 < n) {
                V[(size_t)(a + 1) * n + a] += var * -1;          // (a, a+1) = AV(a,a+1) * A(a+1,a+1)
                V[(size_t)a * n + (a + 1)] = V[(size_t)(a + 1) * n + a];
            }
            previousVariance = var;
        }
        for (uint32_t a = 0; a < n; ++a) {
            c.msr[me.rec[a]].scale2 = V[(size_t)a * n + a];
            c.msr[me.rec[a]].scale3 = (a + 1 < n) ? V[(size_t)(a + 1) * n + a] : 0.0;
        }
    } else {
        for (uint32_t a = 0; a < n; ++a) {
            V[(size_t)a * n + a] = c.msr[me.rec[a]].scale2;
            if (a + 1 < n)
                V[(size_t)(a + 1) * n + a] = V[(size_t)a * n + (a + 1)] = c.msr[me.rec[a]].scale3;
        }
    }
    me.vinv = V;
    return inverse_variance(me.vinv.data(), n, false, c.use_ref);
}

// FormCarttoGeoRotationMatrix (MFN:204-233): d(X,Y,Z)/d(phi,lambda,h) at one position, row-major 3x3
void geo_to_cart_jacobian(const Ellipsoid& e, double lat, double lon, double h, double* R)
{
    const double coslat = std::cos(lat), sinlat = std::sin(lat), coslon = std::cos(lon), sinlon = std::sin(lon);
    const double term1_a = e.a * e.e2, one_minus_esq = 1. - e.e2;
    const double nu = prime_vertical(e, lat);
    const double nu_plus_h = nu + h, nu_1minuse2_plus_h = nu * one_minus_esq + h;
    const double term1_b = term1_a * sinlat * coslat;
    const double term1_c = std::pow(1. - e.e2 * sinlat * sinlat, 1.5);
    R[0] = (term1_b * coslat * coslon / term1_c) - (nu_plus_h * sinlat * coslon);
    R[1] = -nu_plus_h * coslat * sinlon;
    R[2] = coslat * coslon;
    R[3] = (term1_b * coslat * sinlon / term1_c) - (nu_plus_h * sinlat * sinlon);
    R[4] = nu_plus_h * coslat * coslon;
    R[5] = coslat * sinlon;
    R[6] = (term1_b * one_minus_esq * sinlat / term1_c) + (nu_1minuse2_plus_h * coslat);
    R[7] = 0.;
    R[8] = sinlat;
}

// First run of a Y cluster given in latitude / longitude / height (UpdateDesignNormalMeasMatrices_Y ADJ:6281-6325,
// 6384-6420; LoadVarianceMatrix_Y ADJ:4563-4644): the original values go to preAdjMeas, orthometric heights are
// reduced with the station's geoid separation (preAdjCorr), the point becomes Cartesian in place (coordType "XYZ",
// station3 keeps the original type) and the whole variance matrix is propagated V_cart = J V_geo J^T with the
// Jacobians taken at the stations' current positions, then written back (SetGPSVarianceMatrix).
void convert_y_cluster_llh(Ctx& c, Meas& me)
{
    dna_msr_t* m0 = &c.msr[me.first];
    const bool LLH = std::strncmp(m0->coordType, "LLH", 3) == 0, LLh = std::strncmp(m0->coordType, "LLh", 3) == 0;
    if (!LLH && !LLh)
        return;
    const uint32_t n = me.nrows, members = (uint32_t)me.rec.size();
    std::vector<double> V((size_t)n * n, 0.0), J((size_t)n * n, 0.0);
    auto sym = [&](uint32_t r, uint32_t col, double v) { V[(size_t)r * n + col] = V[(size_t)col * n + r] = v; };
    for (uint32_t k = 0; k < members; ++k) {
        dna_msr_t* r = &c.msr[me.rec[k]];
        const dna_stn_t& st = c.stn[r->station1];
        const uint32_t v = 3 * k;
        sym(v, v, r[0].term2);
        sym(v, v + 1, r[1].term2);
        sym(v + 1, v + 1, r[1].term3);
        sym(v, v + 2, r[2].term2);
        sym(v + 1, v + 2, r[2].term3);
        sym(v + 2, v + 2, r[2].term4);
        for (uint32_t q = 0; q < r[0].vectorCount2; ++q) {
            dna_msr_t* cv = r + 3 + 3 * q;
            const uint32_t cc = v + 3 + 3 * q;
            for (int i = 0; i < 3; ++i) {
                sym(v + i, cc, cv[i].term1);
                sym(v + i, cc + 1, cv[i].term2);
                sym(v + i, cc + 2, cv[i].term3);
            }
        }
        double R[9];
        geo_to_cart_jacobian(c.ell, st.currentLatitude, st.currentLongitude, st.currentHeight, R);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                J[(size_t)(v + a) * n + v + b] = R[3 * a + b];
        // the point itself
        const double lat = r[0].term1, lon = r[1].term1;
        double h = r[2].term1;
        for (int q = 0; q < 3; ++q)
            r[q].preAdjMeas = r[q].term1;
        if (LLH && std::fabs(st.geoidSep) > 1.0e-4) {
            r[2].preAdjCorr = st.geoidSep;
            h += r[2].preAdjCorr;
        }
        double x, y, z;
        geo_to_cart(c.ell, lat, lon, h, &x, &y, &z);
        r[0].term1 = x;
        r[1].term1 = y;
        r[2].term1 = z;
        for (int q = 0; q < 3; ++q) {
            std::snprintf(r[q].coordType, sizeof(r[q].coordType), "%s", "XYZ");
            r[q].station3 = LLH ? DNA_LLH_TYPE : DNA_LLh_TYPE;
        }
    }
    // V_cart = J V J^T
    std::vector<double> T((size_t)n * n, 0.0), W((size_t)n * n, 0.0);
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t k = 0; k < n; ++k) {
            const double a = J[(size_t)i * n + k];
            if (a != 0.0)
                for (uint32_t j = 0; j < n; ++j)
                    T[(size_t)i * n + j] += a * V[(size_t)k * n + j];
        }
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) {
            double sum = 0.0;
            for (uint32_t k = 0; k < n; ++k)
                sum += T[(size_t)i * n + k] * J[(size_t)j * n + k];
            W[(size_t)i * n + j] = sum;
        }
    for (uint32_t k = 0; k < members; ++k) {
        dna_msr_t* r = &c.msr[me.rec[k]];
        const uint32_t v = 3 * k;
        r[0].term2 = W[(size_t)v * n + v];
        r[1].term2 = W[(size_t)v * n + v + 1];
        r[1].term3 = W[(size_t)(v + 1) * n + v + 1];
        r[2].term2 = W[(size_t)v * n + v + 2];
        r[2].term3 = W[(size_t)(v + 1) * n + v + 2];
        r[2].term4 = W[(size_t)(v + 2) * n + v + 2];
        for (uint32_t q = 0; q < r[0].vectorCount2; ++q) {
            dna_msr_t* cv = r + 3 + 3 * q;
            const uint32_t cc = v + 3 + 3 * q;
            for (int i = 0; i < 3; ++i) {
                cv[i].term1 = W[(size_t)(v + i) * n + cc];
                cv[i].term2 = W[(size_t)(v + i) * n + cc + 1];
                cv[i].term3 = W[(size_t)(v + i) * n + cc + 2];
            }
        }
    }
}

// LoadVarianceMatrix_X / _Y (ADJ:4312-4450, ADJ:4494-4679) for Cartesian clusters: upper triangle from the records
// (GetGPSVarianceMatrix, MFN:85-123), whole-matrix scalar applied and written back on the first run.
int load_variance_matrix_XY(Ctx& c, Meas& me, bool build)
{
    const uint32_t n = me.nrows, members = (uint32_t)me.rec.size();
    dna_msr_t* m0 = &c.msr[me.first];
    double vS, pS, lS, hS;
    bool scaleMatrix, scalePartial;
    load_variance_scaling(c, *m0, vS, pS, lS, hS, scaleMatrix, scalePartial);
    if (!build)
        scaleMatrix = scalePartial = false;   // the records already hold the scaled matrix (ADJ:4281)
    if (me.type == 'Y' && std::strncmp(m0->coordType, "XYZ", 3) != 0) {
        g_err = "oracle: Y cluster coordinates must be XYZ, LLH or LLh";
        return 3;
    }
    std::vector<double> V((size_t)n * n, 0.0);
    // X clusters apply the whole-matrix scalar while loading (ADJ:4358-4392), Y clusters afterwards and only when no
    // partial scalars are given (ADJ:4646-4647)
    const double onload = (me.type == 'X' && scaleMatrix) ? vS : 1.0;
    auto put = [&](uint32_t r, uint32_t col, double field) {
        V[(size_t)col * n + r] = V[(size_t)r * n + col] = field * onload;
    };
    for (uint32_t k = 0; k < members; ++k) {
        dna_msr_t* r = &c.msr[me.rec[k]];
        const uint32_t v = 3 * k;
        put(v, v, r[0].term2);
        put(v, v + 1, r[1].term2);
        put(v + 1, v + 1, r[1].term3);
        put(v, v + 2, r[2].term2);
        put(v + 1, v + 2, r[2].term3);
        put(v + 2, v + 2, r[2].term4);
        const uint32_t ncov = r[0].vectorCount2;
        for (uint32_t q = 0; q < ncov; ++q) {
            dna_msr_t* cv = r + 3 + 3 * q;
            const uint32_t cc = v + 3 + 3 * q;
            for (int i = 0; i < 3; ++i) {
                put(v + i, cc, cv[i].term1);
                put(v + i, cc + 1, cv[i].term2);
                put(v + i, cc + 2, cv[i].term3);
            }
        }
    }
    if (scalePartial) {
        // ScaleGPSVCV_Cluster (MFN:401-438): to the geographic frame with the Jacobians at the first stations' current
        // positions, scale by sqrt(p), sqrt(l), sqrt(h) (already multiplied by the whole-matrix scalar when both are
        // given, ADJ:4484-4490), back to Cartesian:  V' = (J S J^-1) V (J S J^-1)^T  block by block
        std::vector<double> M((size_t)members * 9);
        for (uint32_t k = 0; k < members; ++k) {
            const dna_stn_t& st = c.stn[c.msr[me.rec[k]].station1];
            double J[9], Ji[9];
            geo_to_cart_jacobian(c.ell, st.currentLatitude, st.currentLongitude, st.currentHeight, J);
            const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
            Ji[0] = (J[4] * J[8] - J[5] * J[7]) / det;
            Ji[1] = (J[2] * J[7] - J[1] * J[8]) / det;
            Ji[2] = (J[1] * J[5] - J[2] * J[4]) / det;
            Ji[3] = (J[5] * J[6] - J[3] * J[8]) / det;
            Ji[4] = (J[0] * J[8] - J[2] * J[6]) / det;
            Ji[5] = (J[2] * J[3] - J[0] * J[5]) / det;
            Ji[6] = (J[3] * J[7] - J[4] * J[6]) / det;
            Ji[7] = (J[1] * J[6] - J[0] * J[7]) / det;
            Ji[8] = (J[0] * J[4] - J[1] * J[3]) / det;
            const double sc[3] = {std::sqrt(pS), std::sqrt(lS), std::sqrt(hS)};
            for (int a = 0; a < 3; ++a)
                for (int b2 = 0; b2 < 3; ++b2) {
                    double sum = 0.0;
                    for (int z = 0; z < 3; ++z)
                        sum += J[3 * a + z] * sc[z] * Ji[3 * z + b2];
                    M[9 * (size_t)k + 3 * a + b2] = sum;
                }
        }
        std::vector<double> W((size_t)n * n, 0.0);
        for (uint32_t ka = 0; ka < members; ++ka)
            for (uint32_t kb = 0; kb < members; ++kb) {
                double T[9];
                for (int x = 0; x < 3; ++x)
                    for (int y = 0; y < 3; ++y) {
                        double sum = 0.0;
                        for (int z = 0; z < 3; ++z)
                            sum += M[9 * (size_t)ka + 3 * x + z] * V[(size_t)(3 * ka + z) * n + 3 * kb + y];
                        T[3 * x + y] = sum;
                    }
                for (int x = 0; x < 3; ++x)
                    for (int y = 0; y < 3; ++y) {
                        double sum = 0.0;
                        for (int z = 0; z < 3; ++z)
                            sum += T[3 * x + z] * M[9 * (size_t)kb + 3 * y + z];
                        W[(size_t)(3 * ka + x) * n + 3 * kb + y] = sum;
                    }
            }
        V.swap(W);
    } else if (me.type == 'Y' && scaleMatrix) {
        for (double& x : V)
            x *= vS;
    }
    if (scaleMatrix || scalePartial) {
        // SetGPSVarianceMatrix: the scaled matrix replaces the record values (ADJ:4425, 4654)
        for (uint32_t k = 0; k < members; ++k) {
            dna_msr_t* r = &c.msr[me.rec[k]];
            const uint32_t v = 3 * k;
            r[0].term2 = V[(size_t)v * n + v];
            r[1].term2 = V[(size_t)v * n + v + 1];
            r[1].term3 = V[(size_t)(v + 1) * n + v + 1];
            r[2].term2 = V[(size_t)v * n + v + 2];
            r[2].term3 = V[(size_t)(v + 1) * n + v + 2];
            r[2].term4 = V[(size_t)(v + 2) * n + v + 2];
            for (uint32_t q = 0; q < r[0].vectorCount2; ++q) {
                dna_msr_t* cv = r + 3 + 3 * q;
                const uint32_t cc = v + 3 + 3 * q;
                for (int i = 0; i < 3; ++i) {
                    cv[i].term1 = V[(size_t)(v + i) * n + cc];
                    cv[i].term2 = V[(size_t)(v + i) * n + cc + 1];
                    cv[i].term3 = V[(size_t)(v + i) * n + cc + 2];
                }
            }
        }
    }
    me.vinv = V;
    return inverse_variance(me.vinv.data(), n, false, c.use_ref);
}

// FillDesignNormalMeasurementsMatrices (ADJ:3888-4055).
//   build=true : first pass — l, partials, V^-1 (with variance scaling write-back), first-run reductions
//   build=false: re-linearise — l (and the partials of the non-GNSS rows; GNSS design never changes, ADJ:5294-5301)
int fill_design_normals(Ctx& c, bool build)
{
    for (Meas& me : c.meas) {
        dna_msr_t* m = &c.msr[me.first];
        uint32_t row = me.row0;
        switch (me.type) {
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
                std::memcpy(&c.vinv[me.g * 9], Vinv.v, sizeof(Vinv.v));
            }
            break;
        }
        case 'X':
        case 'Y': {
            // UpdateDesignNormalMeasMatrices_X (ADJ:6056-6246) / _Y (ADJ:6249-6566)
            if (build && me.type == 'Y')
                convert_y_cluster_llh(c, me);
            for (size_t k = 0; k < me.rec.size(); ++k) {
                dna_msr_t* r = &c.msr[me.rec[k]];
                uint32_t s1 = r->station1 * 3, s2 = r->station2 * 3;
                Row& rw = c.row[row + 3 * k];
                rw.st[0] = r->station1;
                rw.st[1] = r->station2;
                rw.nst = me.type == 'X' ? 2 : 1;
                for (int q = 0; q < 3; ++q) {
                    // clusters that arrived as latitude / longitude / height keep the original values (ADJ:6353-6356)
                    if (build && !(me.type == 'Y' && (r[q].station3 == DNA_LLH_TYPE || r[q].station3 == DNA_LLh_TYPE)))
                        r[q].preAdjMeas = r[q].term1;
                    c.ell_rows[row + 3 * k + q] =
                        me.type == 'X' ? r[q].term1 - (c.est[s2 + q] - c.est[s1 + q]) : r[q].term1 - c.est[s1 + q];
                }
            }
            if (build) {
                int rc = load_variance_matrix_XY(c, me, true);
                if (rc) {
                    if (g_err.empty())
                        g_err = "oracle: GNSS cluster variance matrix inversion failed";
                    return rc;
                }
            }
            break;
        }
        case 'D': {
            // UpdateDesignNormalMeasMatrices_D (ADJ:5082-5240)
            double previousDirection = m->term1;
            for (uint32_t a = 0; a < me.nrows; ++a) {
                dna_msr_t* d = &c.msr[me.rec[a]];
                dna_msr_t angle = c.msr[me.base[a]];   // scratch angle record (angleRec)
                angle.station3 = d->station2;
                if (build) {
                    angle.term1 = d->term1 - previousDirection;
                    if (angle.term1 < 0)
                        angle.term1 += TWO_PI;
                    if (angle.term1 > TWO_PI)
                        angle.term1 -= TWO_PI;
                    angle.preAdjMeas = angle.term1;   // InitialiseMeasurement inside _A
                } else
                    angle.term1 = d->scale1;
                double l = angle_row(c.ell, &angle, c.stn, c.est.data(), angle.station1, angle.station2, angle.station3, build,
                                     c.row[row + a]);
                c.ell_rows[row + a] = l;
                if (build) {
                    d->scale1 = angle.term1;
                    d->preAdjMeas = angle.preAdjMeas;
                    d->preAdjCorr = angle.preAdjCorr;
                    previousDirection = d->term1;
                }
            }
            int rc = load_variance_matrix_D(c, me, build);
            if (rc) {
                g_err = "oracle: direction-set variance matrix inversion failed";
                return rc;
            }
            break;
        }
        default: {
            if (!scalar_row_oracle(c.ell, m, c.stn, c.est.data(), build, c.row[row], &c.ell_rows[row])) {
                g_err = std::string("oracle: measurement type '") + m->measType + "' not restated";
                return 3;
            }
        }
        }
    }
    return 0;
}

// dense A (nrows x 3*|stations|) of one non-G measurement over its unique station list
struct LocalDesign {
    std::vector<uint32_t> stations;
    std::vector<double> A;   // row-major nrows x 3*ns
    uint32_t ns = 0;
};
void local_design(const Ctx& c, const Meas& me, LocalDesign& d)
{
    d.stations.clear();
    auto add = [&](uint32_t s) {
        for (uint32_t t : d.stations)
            if (t == s)
                return;
        d.stations.push_back(s);
    };
    auto find = [&](uint32_t s) {
        for (uint32_t i = 0; i < d.stations.size(); ++i)
            if (d.stations[i] == s)
                return i;
        return 0u;
    };
    const bool cluster_gnss = me.type == 'X' || me.type == 'Y';
    for (uint32_t r = 0; r < me.nrows; ++r) {
        const Row& rw = c.row[me.row0 + (cluster_gnss ? 3 * (r / 3) : r)];
        for (int k = 0; k < rw.nst; ++k)
            add(rw.st[k]);
    }
    d.ns = (uint32_t)d.stations.size();
    d.A.assign((size_t)me.nrows * 3 * d.ns, 0.0);
    for (uint32_t r = 0; r < me.nrows; ++r) {
        double* ar = &d.A[(size_t)r * 3 * d.ns];
        if (cluster_gnss) {
            const Row& rw = c.row[me.row0 + 3 * (r / 3)];
            const int q = r % 3;
            if (me.type == 'X') {
                ar[3 * find(rw.st[0]) + q] += -1.;   // ADJ:5330-5344
                ar[3 * find(rw.st[1]) + q] += 1.;
            } else
                ar[3 * find(rw.st[0]) + q] += 1.;
        } else {
            const Row& rw = c.row[me.row0 + r];
            for (int k = 0; k < rw.nst; ++k)
                for (int q = 0; q < 3; ++q)
                    ar[3 * find(rw.st[k]) + q] += rw.a[3 * k + q];
        }
    }
}
// At V^-1 (3*ns x nrows, row-major) of one non-G measurement: UpdateAtVinv (ADJ:1285-1320), UpdateAtVinv_D
// (ADJ:1328-1356), the X / Y blocks (ADJ:6125-6160, ADJ:6480-6512)
void local_atvinv(const Ctx& c, const Meas& me, const LocalDesign& d, std::vector<double>& AtVinv)
{
    const uint32_t n = me.nrows, w = 3 * d.ns;
    AtVinv.assign((size_t)w * n, 0.0);
    if (me.vinv.empty()) {
        const double variance = 1. / c.msr[me.first].term2;
        for (uint32_t j = 0; j < w; ++j)
            AtVinv[j] = variance * d.A[j];
        return;
    }
    for (uint32_t j = 0; j < w; ++j)
        for (uint32_t r = 0; r < n; ++r) {
            double s = 0.;
            for (uint32_t b = 0; b < n; ++b)
                s += d.A[(size_t)b * w + j] * me.vinv[(size_t)r * n + b];
            AtVinv[(size_t)j * n + r] = s;
        }
}

// UpdateNormals (ADJ:1364-1455): N from the stored At V^-1 / design of every measurement
void update_normals(Ctx& c)
{
    LocalDesign d;
    std::vector<double> AtVinv;
    for (const Meas& me : c.meas) {
        const dna_msr_t* m = &c.msr[me.first];
        if (me.type == 'G') {
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            // UpdateNormals_G (ADJ:1664-1684) through add_normal_3x3_from_atvinv_columns (ADJ:1478-1491):
            // AtVinv[s1.., rows] = -V^-1 ; AtVinv[s2.., rows] = +V^-1
            M3 Vinv;
            std::memcpy(Vinv.v, &c.vinv[me.g * 9], sizeof(Vinv.v));
            for (int col = 0; col < 3; ++col)
                for (int r = 0; r < 3; ++r) {
                    lower_add(c, s2 + r, s2 + col, 1. * Vinv(r, col));
                    lower_add(c, s1 + r, s1 + col, -1. * (-Vinv(r, col)));
                    lower_add(c, s1 + r, s2 + col, 1. * (-Vinv(r, col)));
                    lower_add(c, s2 + r, s1 + col, -1. * Vinv(r, col));
                }
            continue;
        }
        // UpdateNormals_A / _BCEKLMSVZ / _HIJPQR (ADJ:1510-1660), _D (ADJ:1540-1637), _X (ADJ:1687-1787), _Y:
        // N[si, sj] += At V^-1[si, rows] . A[rows, sj] for every pair of the measurement's stations
        local_design(c, me, d);
        local_atvinv(c, me, d, AtVinv);
        const uint32_t n = me.nrows, w = 3 * d.ns;
        for (uint32_t bi = 0; bi < d.ns; ++bi)
            for (uint32_t bj = 0; bj < d.ns; ++bj)
                for (int col = 0; col < 3; ++col)
                    for (int r = 0; r < 3; ++r) {
                        double s = 0.;
                        for (uint32_t q = 0; q < n; ++q)
                            s += AtVinv[(size_t)(3 * bi + r) * n + q] * d.A[(size_t)q * w + 3 * bj + col];
                        lower_add(c, 3 * d.stations[bi] + r, 3 * d.stations[bj] + col, s);
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
    LocalDesign d;
    std::vector<double> AtVinv;
    for (const Meas& me : c.meas) {
        const dna_msr_t* m = &c.msr[me.first];
        const uint32_t row = me.row0;
        if (me.type == 'G') {
            uint32_t s1 = m->station1 * 3, s2 = m->station2 * 3;
            const double* V = &c.vinv[me.g * 9];
            for (int r = 0; r < 3; ++r) {
                double t = 0.;
                for (int k = 0; k < 3; ++k)
                    t += V[k * 3 + r] * c.ell_rows[row + k];
                c.w[s1 + r] += -t;
                c.w[s2 + r] += t;
            }
            continue;
        }
        local_design(c, me, d);
        local_atvinv(c, me, d, AtVinv);
        for (uint32_t b = 0; b < d.ns; ++b)
            for (int r = 0; r < 3; ++r) {
                double t = 0.;
                for (uint32_t q = 0; q < me.nrows; ++q)
                    t += AtVinv[(size_t)(3 * b + r) * me.nrows + q] * c.ell_rows[row + q];
                c.w[3 * d.stations[b] + r] += t;
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

// ComputeStatistics (ADJ:7116-7148) over one block / the whole network: c.N holds the rigorous variances (the inverse),
// c.ell_rows the measured - computed values at the final estimates.  Adds to chi / outliers, writes the records.
int compute_statistics(Ctx& c, double critical, double& chi, uint32_t& outliers)
{
    dna_msr_t* msr = c.msr;
    dna_stn_t* stn = c.stn;
    int rc = 0;
        auto Q = [&](uint32_t i, uint32_t j) { return i >= j ? c.N[pidx(c.n, i, j)] : c.N[pidx(c.n, j, i)]; };
        // ComputePrecisionAdjMsrs_A / _BCEKLMSVZ / _HIJPQR (ADJ:7880-8003): a Q a^T over the row's stations
        auto row_precision = [&](const Row& rw) {
            double prec = 0.;
            for (int bs = 0; bs < rw.nst; ++bs)
                for (int i = 0; i < 3; ++i) {
                    double part = 0.;
                    for (int bj = 0; bj < rw.nst; ++bj)
                        for (int k = 0; k < 3; ++k)
                            part += rw.a[3 * bj + k] * Q(3 * rw.st[bj] + k, 3 * rw.st[bs] + i);
                    prec += part * rw.a[3 * bs + i];
                }
            return prec;
        };
        for (Meas& me : c.meas) {
            dna_msr_t* m = &msr[me.first];
            const uint32_t row = me.row0;
            switch (me.type) {
            case 'G':
            case 'X':
            case 'Y': {
                const size_t members = me.type == 'G' ? 1 : me.rec.size();
                for (size_t k = 0; k < members; ++k) {
                    dna_msr_t* r = me.type == 'G' ? m : &msr[me.rec[k]];
                    uint32_t s1 = r->station1 * 3, s2 = r->station2 * 3;
                    double p6[6];
                    if (me.type == 'Y') {
                        // ComputePrecisionAdjMsrs_Y (ADJ:8035-8060)
                        int q = 0;
                        for (int i = 0; i < 3; ++i)
                            for (int j = i; j < 3; ++j)
                                p6[q++] = Q(s1 + i, s1 + j);
                    } else
                        precision_adjusted_gnss_bsl(c, s1, s2, p6);   // ComputePrecisionAdjMsrs_GX (ADJ:8006-8032)
                    // UpdateMsrRecords_GXY (ADJ:8152-8184): XX row+0, YY row+3, ZZ row+5
                    const uint32_t rr = row + 3 * (uint32_t)k;
                    update_msr_record(r[0], c.ell_rows[rr + 0], p6[0], r[0].term2, critical, outliers);
                    update_msr_record(r[1], c.ell_rows[rr + 1], p6[3], r[1].term3, critical, outliers);
                    update_msr_record(r[2], c.ell_rows[rr + 2], p6[5], r[2].term4, critical, outliers);
                }
                if (me.type == 'G') {
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
                } else {
                    // ComputeChiSquare_XY (ADJ:8552-8577): r^T V^-1 r with V^-1 from the records as they stand
                    rc = load_variance_matrix_XY(c, me, false);
                    if (rc)
                        return rc;
                    const uint32_t n = me.nrows;
                    double cs = 0.;
                    for (uint32_t j = 0; j < n; ++j) {
                        double t = 0.;
                        for (uint32_t i = 0; i < n; ++i)
                            t += c.ell_rows[row + i] * me.vinv[(size_t)j * n + i];
                        cs += t * c.ell_rows[row + j];
                    }
                    chi += cs;
                }
                break;
            }
            case 'D': {
                // ComputePrecisionAdjMsrs_D (ADJ:7912-7946), UpdateMsrRecords_D (ADJ:8120-8149), ComputeChiSquare_D (ADJ:8440-8469)
                for (uint32_t a = 0; a < me.nrows; ++a) {
                    dna_msr_t& d = msr[me.rec[a]];
                    const double prec = row_precision(c.row[row + a]);
                    update_msr_record(d, c.ell_rows[row + a], prec, d.scale2, critical, outliers);
                    d.measAdj = d.scale1 + d.measCorr;   // ADJ:8194-8199
                    if (d.measAdj > TWO_PI)
                        d.measAdj -= TWO_PI;
                    d.measAdj += d.preAdjCorr;           // ADJ:8259-8268
                    chi += c.ell_rows[row + a] * c.ell_rows[row + a] / d.scale2;
                }
                break;
            }
            default: {
                const Row& rw = c.row[row];
                const double prec = row_precision(rw);
                update_msr_record(*m, c.ell_rows[row], prec, m->term2, critical, outliers);
                const double* p1 = &c.est[3 * (size_t)m->station1];
                const double* p2 = &c.est[3 * (size_t)m->station2];
                const dna_stn_t& st1 = stn[m->station1];
                switch (m->measType) {   // ADJ:8205-8271
                case 'E':
                    m->measAdj = ell_chord_to_arc(c.ell, m->measAdj, p1, p2, st1.currentLatitude, st1.currentLongitude,
                                                  stn[m->station2].currentLatitude);
                    break;
                case 'M':
                    m->measAdj = ell_chord_to_msl_arc(c.ell, m->measAdj, st1.currentLatitude, stn[m->station2].currentLatitude,
                                                      st1.geoidSep, stn[m->station2].geoidSep);
                    break;
                case 'H':
                case 'L':
                    m->measAdj -= m->preAdjCorr;
                    break;
                case 'A':
                case 'I':
                case 'J':
                case 'K':
                case 'Z':
                    m->measAdj += m->preAdjCorr;
                    break;
                case 'V':
                    m->measAdj -= m->preAdjCorr;
                    break;
                }
                chi += c.ell_rows[row] * c.ell_rows[row] / m->term2;   // ADJ:8430-8437
            }
            }
        }
    return rc;
}

// ComputeGlobalPelzer (ADJ:8302-8427): tally over one block
void pelzer_tally(Ctx& c, double& sum, uint32_t& num)
{
    dna_msr_t* msr = c.msr;
    auto tally = [&](dna_msr_t& r, double limit) {
        if (r.PelzerRel > 0. && r.PelzerRel < limit) {
            sum += (r.PelzerRel * r.PelzerRel - 1.);
            num++;
        } else
            r.PelzerRel = UNRELIABLE;
    };
    for (Meas& me : c.meas) {
        dna_msr_t* m = &msr[me.first];
        switch (me.type) {
        case 'G':
            for (int k = 0; k < 3; ++k)
                tally(m[k], UNRELIABLE);
            break;
        case 'X':
        case 'Y':
            for (uint64_t r0 : me.rec)
                for (int k = 0; k < 3; ++k)
                    tally(msr[r0 + k], UNRELIABLE);
            break;
        case 'D':
            for (uint64_t r0 : me.rec)
                tally(msr[r0], UNRELIABLE);   // ComputeGlobalPelzer_D (ADJ:8362-8393)
            break;
        default:
            tally(*m, STABLE_LIMIT);          // ADJ:8339-8345
        }
    }
}

// ---- phased adjustment (AdjustPhased, ADJ:2579-2670) --------------------------------------------------------------
// A network segmented into a chain of blocks (dnasegment's .seg): block b holds its inner stations ISL(b) and its
// junction stations JSL(b), the stations that appear again in block b+1 (seg_file.cpp:432-486).  Every block is a small
// dense network of its own here: copies of its station and measurement records with block-local station numbers, so that
// the simultaneous-mode restatements above (design rows, normals, statistics) serve unchanged per block.
uint64_t record_span(const dna_msr_t* msr, uint64_t nmsr, uint64_t i)
{
    const dna_msr_t& m = msr[i];
    switch (m.measType) {
    case 'G':
        return 3;
    case 'X':
    case 'Y': {
        uint64_t j = i;
        for (uint32_t k = 0; k < m.vectorCount1 && j < nmsr; ++k)
            j += 3 + 3ull * msr[j].vectorCount2;
        return j - i;
    }
    case 'D':
        return m.vectorCount1 ? m.vectorCount1 : 1;
    default:
        return 1;
    }
}

// stations a measurement's rows refer to (the station lists dnasegment builds the blocks from)
void stations_of(const dna_msr_t* msr, const Meas& me, std::vector<uint32_t>& out)
{
    out.clear();
    auto touch = [&](const dna_msr_t& r) {
        out.push_back(r.station1);
        if (me.type != 'Y' && !std::strchr("HRIJPQ", me.type))
            out.push_back(r.station2);
        if (me.type == 'A')
            out.push_back(r.station3);
    };
    touch(msr[me.first]);
    for (uint64_t r : me.rec)
        touch(msr[r]);
    for (uint64_t r : me.base)
        touch(msr[r]);
}

struct Block {
    std::vector<uint32_t> stations;           // v_parameterStationList_: inner stations, then junction stations (global indices)
    uint32_t n_inner = 0;
    std::vector<uint32_t> jprev;              // local index in this block of every station of JSL(b-1), in JSL(b-1) order
    std::vector<uint8_t> first_fwd;           // per local station: first appearance in a forward pass (seg_file.cpp:432-486)
    std::vector<dna_stn_t> stn;               // block-local copies
    std::vector<dna_msr_t> msr;
    std::vector<uint64_t> src;                // source index of every copied record
    Ctx c;
    uint32_t n = 0;
    std::vector<double> orig;                 // v_originalStations_
    std::vector<double> corr_final;           // corrections of the rigorous solution of this iteration
    std::vector<double> jvar_fwd, jest_fwd;   // v_junctionVariancesFwd_ (inverted, full column-major), v_junctionEstimatesFwd_ over JSL(b)
    std::vector<double> Q;                    // v_rigorousVariances_ (packed)
    explicit Block(const oracle_opts* o) : c(o) {}
    uint32_t njunction() const { return (uint32_t)stations.size() - n_inner; }
};

// measurement part of the normals and of At V^-1 l at the block's current estimates (RebuildNormals ADJ:2721-2753 without
// the parameter-station step; v_normalsR_ "contributions from all apriori measurement variances")
void block_measurement_system(Block& B)
{
    std::fill(B.c.N.begin(), B.c.N.end(), 0.0);
    update_normals(B.c);
    weighted_rhs(B.c);
}

// AddConstraintStationstoNormalsForward / Reverse / Combine (ADJ:1884-2002): sign +1 adds, -1 removes
int block_constraints(Block& B, int mode /*0 forward, 1 reverse, 2 combine*/)
{
    for (uint32_t s = 0; s < B.stations.size(); ++s) {
        const bool first_rev = s < B.n_inner;   // last block the station appears in = the block it is inner to
        double sign = 1.0;
        if (mode == 0 && !B.first_fwd[s])
            continue;
        if (mode == 1 && !first_rev)
            continue;
        if (mode == 2) {
            if (B.first_fwd[s])
                continue;
            sign = -1.0;                         // carried in from the forward pass already: applied twice otherwise
        }
        M3 V;
        int rc = constraint_block(B.c, B.stn[s], V);
        if (rc)
            return rc;
        for (int col = 0; col < 3; ++col)
            for (int r = col; r < 3; ++r)
                B.c.N[pidx(B.n, s * 3 + r, s * 3 + col)] += sign * V(r, col);
    }
    return 0;
}

// Junction stations as pseudo measurements (CarryStnEstimatesandVariances{Forward,Reverse,Combine}, ADJ:998-1281,
// 3196-3333): the inverted junction variance matrix J (full, 3j x 3j, over `idx`) is added to the normals and
// J (carried estimates - current estimates) to At V^-1 l  (the reference grows AtVinv / measMinusComp by the pseudo rows
// and multiplies; the product is this sum).
void add_junction_pseudo_measurements(Block& B, const std::vector<uint32_t>& idx, const std::vector<double>& J,
                                      const std::vector<double>& jest)
{
    const uint32_t nj = 3 * (uint32_t)idx.size();
    std::vector<double> d(nj);
    for (uint32_t a = 0; a < idx.size(); ++a)
        for (int k = 0; k < 3; ++k)
            d[3 * a + k] = jest[3 * a + k] - B.c.est[3 * idx[a] + k];
    for (uint32_t a = 0; a < nj; ++a) {
        const uint32_t ra = 3 * idx[a / 3] + a % 3;
        double t = 0.;
        for (uint32_t b = 0; b < nj; ++b) {
            const uint32_t rb = 3 * idx[b / 3] + b % 3;
            const double v = J[(size_t)b * nj + a];
            if (ra >= rb)
                B.c.N[pidx(B.n, ra, rb)] += v;
            t += v * d[b];
        }
        B.c.w[ra] += t;
    }
}

// Solve (ADJ:6586-6667) on a block whose normals and At V^-1 l are complete; estimates += corrections
int block_solve(Block& B, double* t_inv)
{
    Ctx& c = B.c;
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
        g_err = "Matrix inversion failed, the matrix is singular.";
        return rc;
    }
    if (c.o->scale_normals_to_unity)
        for (uint32_t j = 0; j < c.n; ++j)
            for (uint32_t i = j; i < c.n; ++i)
                c.N[pidx(c.n, i, j)] *= sdiag[i] * sdiag[j];
    sym_packed_mv(c.N, c.n, c.w.data(), c.corr.data(), c.use_ref);
    for (uint32_t k = 0; k < c.n; ++k)
        c.est[k] += c.corr[k];
    return 0;
}

// variances and estimates of the stations `idx` of an adjusted block -> inverted junction variance matrix + estimates
int junction_carry(const Block& B, const std::vector<uint32_t>& idx, bool use_ref, std::vector<double>& J, std::vector<double>& jest)
{
    const uint32_t nj = 3 * (uint32_t)idx.size();
    J.assign((size_t)nj * nj, 0.0);
    jest.resize(nj);
    auto Q = [&](uint32_t i, uint32_t j) { return i >= j ? B.c.N[pidx(B.n, i, j)] : B.c.N[pidx(B.n, j, i)]; };
    for (uint32_t a = 0; a < nj; ++a) {
        const uint32_t ra = 3 * idx[a / 3] + a % 3;
        jest[a] = B.c.est[ra];
        for (uint32_t b = 0; b < nj; ++b)
            J[(size_t)b * nj + a] = Q(ra, 3 * idx[b / 3] + b % 3);
    }
    if (nj == 0)
        return 0;
    int rc = inverse_variance(J.data(), nj, false, use_ref);   // FormInverseVarianceMatrix (ADJ:1053, 1201)
    if (rc)
        g_err = "Matrix inversion failed, the junction station variance matrix is singular.";
    return rc;
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
    {
        size_t g = 0;
        for (Meas& me : c.meas) {
            me.row0 = c.rows;
            c.rows += me.nrows;
            if (me.type == 'G')
                me.g = g++;
        }
    }

    // PopulateEstimatedStationMatrix (ADJ:632-693)
    c.est.resize(c.n);
    // stations without a measurement that takes part are not in the reference's station lists (RemoveInvalidStations,
    // network_data_loader.cpp:286-300), so they are not unknowns (ADJ:632-693 runs over the lists)
    std::vector<uint8_t> used(nstn, 0);
    auto touch = [&](const dna_msr_t& r, char type) {
        used[r.station1] = 1;
        if (type != 'Y' && !std::strchr("HRIJPQ", type))
            used[r.station2] = 1;
        if (type == 'A')
            used[r.station3] = 1;
    };
    for (const Meas& me : c.meas) {
        touch(msr[me.first], me.type);
        for (uint64_t r : me.rec)
            touch(msr[r], me.type);
        for (uint64_t r : me.base)
            touch(msr[r], me.type);
    }
    uint32_t unknownParams = 0;
    for (uint32_t s = 0; s < nstn; ++s) {
        geo_to_cart(c.ell, stn[s].currentLatitude, stn[s].currentLongitude, stn[s].currentHeight, &c.est[3 * s],
                    &c.est[3 * s + 1], &c.est[3 * s + 2]);
        if (!used[s])
            continue;
        unknownParams += 3;
        for (int k = 0; k < 3; ++k)
            if (stn[s].stationConst[k] == 'C')
                unknownParams--;
    }
    c.N.assign(psize(c.n), 0.0);
    c.ell_rows.assign(c.rows, 0.0);
    c.vinv.assign(c.cml.size() * 9, 0.0);
    c.row.assign(c.rows, Row());
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
    rc = compute_statistics(c, critical, chi, outliers);
    if (rc)
        return rc;
    res->chi_squared = chi;
    res->measurement_params = c.rows;
    res->unknown_params = unknownParams;
    res->dof = (int64_t)c.rows - (int64_t)unknownParams;  // ADJ:6856
    res->sigma_zero = res->dof != 0 ? chi / (double)res->dof : 0.;
    res->outliers = outliers;
    {
        double sum = 0.;
        uint32_t num = 0;
        pelzer_tally(c, sum, num);
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

/* AdjustPhased (ADJ:2579-2670): forward pass (ADJ:2756-2852), reverse + combination pass (ADJ:3461-3590) per iteration. */
int oracle_adjust_phased(const oracle_opts* opts, dna_stn_t* stn, uint32_t nstn, dna_msr_t* msr, uint64_t nmsr, uint32_t nblocks,
                         const uint32_t* isl_off, const uint32_t* isl, double* est_xyz, double* stn_vcv, int32_t want_block,
                         uint32_t* block_nstn, uint32_t* block_stations, double* block_vcv, oracle_result* res)
{
    g_err.clear();
    std::memset(res, 0, sizeof(*res));
    const bool use_ref = opts->use_ref && g_ref.h;
    if (use_ref && g_ref.set_threads && opts->threads > 0)
        g_ref.set_threads(opts->threads);
    res->used_ref = use_ref ? 1 : 0;
    double t0 = now_s();
    double conf = opts->confidence_interval * 0.01;
    conf += (1.0 - conf) / 2.0;
    const double critical = norm_quantile(conf);
    res->critical_value = critical;
    if (nblocks == 0) {
        g_err = "oracle: no blocks";
        return 2;
    }

    // ---- block lists: ISL given; CML(b) = the measurements whose first-eliminated station is inner to b;
    //      JSL(b) = (stations of CML(b) + JSL(b-1)) - ISL(b)   (what dnasegment writes to the .seg file)
    std::vector<int32_t> block_of(nstn, -1);
    for (uint32_t b = 0; b < nblocks; ++b)
        for (uint32_t k = isl_off[b]; k < isl_off[b + 1]; ++k) {
            if (isl[k] >= nstn || block_of[isl[k]] >= 0) {
                g_err = "oracle: bad inner station list";
                return 2;
            }
            block_of[isl[k]] = (int32_t)b;
        }
    Ctx g(opts);   // the whole list, only to split it into measurements
    g.stn = stn;
    g.nstn = nstn;
    g.msr = msr;
    g.nmsr = nmsr;
    int rc = build_cml(g);
    if (rc)
        return rc;
    std::vector<std::vector<size_t>> cml(nblocks);
    std::vector<uint8_t> used(nstn, 0);
    std::vector<uint32_t> touched;
    uint32_t total_rows = 0;
    for (size_t k = 0; k < g.meas.size(); ++k) {
        stations_of(msr, g.meas[k], touched);
        int32_t b = (int32_t)nblocks;
        for (uint32_t s : touched) {
            if (s >= nstn || block_of[s] < 0) {
                g_err = "oracle: measurement refers to a station outside every block";
                return 2;
            }
            used[s] = 1;
            b = std::min(b, block_of[s]);
        }
        cml[b].push_back(k);
        total_rows += g.meas[k].nrows;
    }
    uint32_t unknownParams = 0;
    for (uint32_t s = 0; s < nstn; ++s) {
        if (!used[s])
            continue;
        unknownParams += 3;
        for (int k = 0; k < 3; ++k)
            if (stn[s].stationConst[k] == 'C')
                unknownParams--;
    }

    std::vector<std::unique_ptr<Block>> blocks;
    std::vector<int32_t> local(nstn, -1);
    std::vector<uint32_t> jprev_global;   // JSL(b-1), global indices
    for (uint32_t b = 0; b < nblocks; ++b) {
        blocks.emplace_back(new Block(opts));
        Block& B = *blocks.back();
        B.stations.assign(isl + isl_off[b], isl + isl_off[b + 1]);
        B.n_inner = (uint32_t)B.stations.size();
        std::vector<uint32_t> junction;
        for (size_t k : cml[b]) {
            stations_of(msr, g.meas[k], touched);
            for (uint32_t s : touched)
                if (block_of[s] != (int32_t)b)
                    junction.push_back(s);
        }
        for (uint32_t s : jprev_global)
            if (block_of[s] != (int32_t)b)
                junction.push_back(s);
        std::sort(junction.begin(), junction.end());
        junction.erase(std::unique(junction.begin(), junction.end()), junction.end());
        for (uint32_t s : junction)
            if (block_of[s] < (int32_t)b) {
                g_err = "oracle: the blocks do not form a chain (a junction station belongs to an earlier block)";
                return 2;
            }
        if (b + 1 == nblocks && !junction.empty()) {
            g_err = "oracle: the last block has junction stations";
            return 2;
        }
        B.stations.insert(B.stations.end(), junction.begin(), junction.end());
        for (uint32_t i = 0; i < B.stations.size(); ++i)
            local[B.stations[i]] = (int32_t)i;
        B.first_fwd.assign(B.stations.size(), 1);
        for (uint32_t s : jprev_global) {
            B.jprev.push_back((uint32_t)local[s]);
            B.first_fwd[local[s]] = 0;
        }
        // block-local copies of the records
        B.stn.resize(B.stations.size());
        for (uint32_t i = 0; i < B.stations.size(); ++i)
            B.stn[i] = stn[B.stations[i]];
        for (size_t k : cml[b]) {
            const Meas& me = g.meas[k];
            const uint64_t span = record_span(msr, nmsr, me.first);
            for (uint64_t r = me.first; r < me.first + span && r < nmsr; ++r) {
                dna_msr_t rec = msr[r];
                auto renumber = [&](uint32_t& st) {
                    if (st < nstn && local[st] >= 0)
                        st = (uint32_t)local[st];
                };
                renumber(rec.station1);
                if (me.type != 'Y' && !std::strchr("HRIJPQ", me.type))
                    renumber(rec.station2);
                if (me.type == 'A')
                    renumber(rec.station3);
                B.msr.push_back(rec);
                B.src.push_back(r);
            }
        }
        for (uint32_t s : B.stations)
            local[s] = -1;
        jprev_global = junction;

        // PrepareAdjustmentBlock (ADJ:2873-3003): design rows, variance matrices, first-run reductions
        Ctx& c = B.c;
        c.stn = B.stn.data();
        c.nstn = (uint32_t)B.stn.size();
        c.msr = B.msr.data();
        c.nmsr = B.msr.size();
        c.n = B.n = 3 * c.nstn;
        c.use_ref = use_ref;
        rc = build_cml(c);
        if (rc)
            return rc;
        size_t gi = 0;
        for (Meas& me : c.meas) {
            me.row0 = c.rows;
            c.rows += me.nrows;
            if (me.type == 'G')
                me.g = gi++;
        }
        c.est.resize(c.n);
        for (uint32_t i = 0; i < c.nstn; ++i)
            geo_to_cart(c.ell, B.stn[i].currentLatitude, B.stn[i].currentLongitude, B.stn[i].currentHeight, &c.est[3 * i],
                        &c.est[3 * i + 1], &c.est[3 * i + 2]);
        B.orig = c.est;
        c.N.assign(psize(c.n), 0.0);
        c.ell_rows.assign(c.rows, 0.0);
        c.vinv.assign(c.cml.size() * 9, 0.0);
        c.row.assign(c.rows, Row());
        c.corr.assign(c.n, 0.0);
        c.w.assign(c.n, 0.0);
        rc = fill_design_normals(c, true);
        if (rc)
            return rc;
    }
    res->seconds_prepare = now_s() - t0;

    const uint32_t NB = nblocks;
    // UpdateAdjustment (ADJ:473-627): geographic coordinates of the station records from the blocks' estimates, in block
    // order (UpdateGeographicCoordsPhased, ADJ:8711-8731: a junction station ends up with the values of the last block
    // that holds it), for the blocks with local-frame measurements, the last block, or every block once the iteration
    // has stopped; then the measured - computed values at the new estimates
    auto update_adjustment = [&](bool iterate) -> int {
        for (uint32_t b = 0; b < NB; ++b) {
            Block& B = *blocks[b];
            Ctx& c = B.c;
            if (c.non_gps || b + 1 == NB || !iterate)
                for (uint32_t i = 0; i < c.nstn; ++i) {
                    dna_stn_t& st = stn[B.stations[i]];
                    cart_to_geo(c.ell, c.est[3 * i], c.est[3 * i + 1], c.est[3 * i + 2], &st.currentLatitude,
                                &st.currentLongitude, &st.currentHeight);
                }
        }
        for (uint32_t b = 0; b < NB; ++b) {
            Block& B = *blocks[b];
            for (uint32_t i = 0; i < B.c.nstn; ++i) {
                const dna_stn_t& st = stn[B.stations[i]];
                B.stn[i].currentLatitude = st.currentLatitude;
                B.stn[i].currentLongitude = st.currentLongitude;
                B.stn[i].currentHeight = st.currentHeight;
            }
            int r2 = fill_design_normals(B.c, false);
            if (r2)
                return r2;
        }
        return 0;
    };
    double maxCorr = 0.;
    uint32_t iter = 0;
    std::vector<double> jvar_rev, jest_rev, NR, wR, est_fwd_last;
    for (uint32_t it = 0; it < opts->max_iterations; ++it) {
        ++iter;
        maxCorr = 0.;
        auto track = [&](const std::vector<double>& corr) {   // compute_maximum_value (MATC:1532-1555)
            size_t mr = 0;
            for (size_t k = 0; k < corr.size(); ++k)
                if (std::fabs(corr[k]) > std::fabs(corr[mr]))
                    mr = k;
            if (!corr.empty() && std::fabs(corr[mr]) > std::fabs(maxCorr))
                maxCorr = corr[mr];
        };
        double ts = now_s();
        // ---- AdjustPhasedForward (ADJ:2756-2852)
        for (uint32_t b = 0; b < NB; ++b) {
            Block& B = *blocks[b];
            B.c.est = B.orig;
            block_measurement_system(B);
            rc = block_constraints(B, 0);
            if (rc)
                return rc;
            if (b > 0)   // CarryStnEstimatesandVariancesForward (ADJ:998-1128) left these with the previous block
                add_junction_pseudo_measurements(B, B.jprev, blocks[b - 1]->jvar_fwd, blocks[b - 1]->jest_fwd);
            rc = block_solve(B, &res->seconds_inverse);   // SolveTry + UpdateEstimatesForward (ADJ:3022-3060)
            if (rc)
                return rc;
            if (b + 1 == NB) {
                // the last block is rigorous after the forward pass
                B.Q = B.c.N;
                B.corr_final = B.c.corr;
                est_fwd_last = B.c.est;
                track(B.c.corr);
            } else {
                std::vector<uint32_t> jidx(B.njunction());
                for (uint32_t i = 0; i < jidx.size(); ++i)
                    jidx[i] = B.n_inner + i;
                rc = junction_carry(B, jidx, use_ref, B.jvar_fwd, B.jest_fwd);
                if (rc)
                    return rc;
            }
        }
        // ---- AdjustPhasedReverseCombine (ADJ:3461-3590); a single block is rigorous already
        for (uint32_t bb = 0; NB > 1 && bb < NB; ++bb) {
            const uint32_t b = NB - 1 - bb;
            Block& B = *blocks[b];
            // PrepareAdjustmentReverse (ADJ:3112-3167) / CarryReverseJunctions (ADJ:3833-3883): original coordinates,
            // the measurement normals, junction stations of the block adjusted before, parameter stations (reverse)
            B.c.est = B.orig;
            block_measurement_system(B);
            if (b + 1 < NB) {
                std::vector<uint32_t> jidx(B.njunction());
                for (uint32_t i = 0; i < jidx.size(); ++i)
                    jidx[i] = B.n_inner + i;
                add_junction_pseudo_measurements(B, jidx, jvar_rev, jest_rev);
            }
            rc = block_constraints(B, 1);
            if (rc)
                return rc;
            const bool combine = b > 0 && b + 1 < NB;   // CombineRequired
            if (combine) {   // BackupNormals (ADJ:3170-3192)
                NR = B.c.N;
                wR = B.c.w;
            }
            rc = block_solve(B, &res->seconds_inverse);   // reverse, in isolation (rigorous for the first block)
            if (rc)
                return rc;
            if (b > 0) {
                // CarryStnEstimatesandVariancesReverse (ADJ:1133-1281): the junction stations shared with block b-1
                rc = junction_carry(B, B.jprev, use_ref, jvar_rev, jest_rev);
                if (rc)
                    return rc;
            }
            if (combine) {
                // PrepareAdjustmentCombine (ADJ:3336-3396), CarryStnEstimatesandVariancesCombine (ADJ:3196-3333)
                B.c.est = B.orig;
                B.c.N = NR;
                B.c.w = wR;
                add_junction_pseudo_measurements(B, B.jprev, blocks[b - 1]->jvar_fwd, blocks[b - 1]->jest_fwd);
                rc = block_constraints(B, 2);
                if (rc)
                    return rc;
                rc = block_solve(B, &res->seconds_inverse);
                if (rc)
                    return rc;
            }
            // UpdateEstimatesFinal (ADJ:3744-3830)
            if (b + 1 == NB) {
                B.c.est = est_fwd_last;   // the rigorous forward solution of the last block stands
                continue;
            }
            B.Q = B.c.N;
            B.corr_final = B.c.corr;
            track(B.c.corr);
        }
        res->seconds_solve += now_s() - ts;
        // rigorous estimates become the original ones of the next iteration (ADJ:3815, 521-523)
        for (auto& pb : blocks)
            pb->orig = pb->c.est;
        const bool iterate = std::fabs(maxCorr) > opts->iteration_threshold;
        if (!iterate)
            break;
        // UpdateAdjustment(true) (ADJ:473-627): geographic coordinates where local-frame measurements need them, new
        // measured - computed; the normals are rebuilt at the start of the next pass
        rc = update_adjustment(true);
        if (rc)
            return rc;
    }
    res->iterations = iter;
    res->max_corr = maxCorr;
    res->converged = std::fabs(maxCorr) <= opts->iteration_threshold;

    // ---- GenerateStatistics (ADJ:6802-6841): UpdateAdjustment(false), then per block with its rigorous variances
    uint32_t outliers = 0;
    double chi = 0., psum = 0.;
    uint32_t pnum = 0;
    rc = update_adjustment(false);
    if (rc)
        return rc;
    for (uint32_t b = 0; b < NB; ++b) {
        Block& B = *blocks[b];
        Ctx& c = B.c;
        c.N = B.Q;
        rc = compute_statistics(c, critical, chi, outliers);
        if (rc)
            return rc;
        pelzer_tally(c, psum, pnum);
        // results back into the caller's records (station numbers restored)
        for (size_t r = 0; r < B.msr.size(); ++r) {
            dna_msr_t out = B.msr[r];
            const dna_msr_t& in = msr[B.src[r]];
            out.station1 = in.station1;
            out.station2 = in.station2;
            if (c.msr[r].measType == 'A')
                out.station3 = in.station3;
            msr[B.src[r]] = out;
        }
        auto Q = [&](uint32_t i, uint32_t j) { return i >= j ? B.Q[pidx(B.n, i, j)] : B.Q[pidx(B.n, j, i)]; };
        for (uint32_t i = 0; i < B.n_inner; ++i) {
            const uint32_t s = B.stations[i];
            if (est_xyz)
                for (int k = 0; k < 3; ++k)
                    est_xyz[3 * (size_t)s + k] = c.est[3 * i + k];
            if (stn_vcv)
                for (int r = 0; r < 3; ++r)
                    for (int k = 0; k < 3; ++k)
                        stn_vcv[9 * (size_t)s + 3 * r + k] = Q(3 * i + r, 3 * i + k);
        }
        if ((int32_t)b == want_block) {
            if (block_nstn)
                *block_nstn = c.nstn;
            if (block_stations)
                std::memcpy(block_stations, B.stations.data(), c.nstn * sizeof(uint32_t));
            if (block_vcv)
                for (uint32_t j = 0; j < B.n; ++j)
                    for (uint32_t i = 0; i < B.n; ++i)
                        block_vcv[(size_t)j * B.n + i] = Q(i, j);
        }
        std::vector<double>().swap(B.Q);
        std::vector<double>().swap(c.N);
    }
    res->chi_squared = chi;
    res->measurement_params = total_rows;
    res->unknown_params = unknownParams;
    res->dof = (int64_t)total_rows - (int64_t)unknownParams;
    res->sigma_zero = res->dof != 0 ? chi / (double)res->dof : 0.;
    res->outliers = outliers;
    res->global_pelzer = pnum > 0 ? std::sqrt(psum / pnum) : UNRELIABLE;
    return 0;
}

}  // extern "C"
