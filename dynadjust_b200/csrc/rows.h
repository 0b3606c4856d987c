// rows.h — design rows of every measurement type other than the single GNSS baseline, shared by the CUDA kernels
// (kernels_rows.cu), the host-side first-run reductions in engine.cpp and the CPU stand-in build under tests/hostsim.
//
// A *row* is one scalar observation equation: residual l = measured - computed, partial derivatives with respect to
// the Cartesian coordinates of one, two or three stations, and a variance.  Rows are either independent (types
// A B C E H I J K L M P Q R S V Z: weight 1/variance) or members of a *cluster* that carries a full inverse
// variance matrix (D direction sets -> derived angles; X baseline clusters; Y point clusters).
//
// Reference semantics per type (dynadjust/dynadjust/dnaadjust/dnaadjust.cpp = ADJ, include/functions/
// dnatemplategeodesyfuncs.hpp = GEO):
//   A  ADJ:4754-4910   B/K ADJ:4913-5014   C/E/M ADJ:5017-5079, 5242-5281, 5398-5428   S ADJ:5437-5493
//   V  ADJ:5504-5601   Z ADJ:5613-5710     L ADJ:5717-5784   H/R ADJ:5969-6053   I/P ADJ:5786-5914   J/Q ADJ:5816-5966
//   D  ADJ:5082-5240 (+ LoadVarianceMatrix_D ADJ:4059-4188)   X ADJ:6056-6246   Y ADJ:6249-6566
// The derivatives are written here in vector form (unit east / north / up vectors of the instrument station) —
// algebraically the reference's expanded expressions (e.g. cos^2(az)/n^2 = 1/(e^2+n^2)).
#pragma once
#include <cstdint>

#include "../../include/dna_records.h"
#include "geodesy.h"

namespace gadj {

constexpr double kTwoPi = kPi + kPi;
constexpr double kHalfPi = kPi / 2.0;
constexpr double kDeflectionEps = 0.0001 * (kPi / 180.0 / 3600.0);   // E4_SEC_DEFLECTION, dnaconsts.hpp:110

// per design row (host-built plan, read-only on the device)
struct RowDesc {
    uint32_t rec;        // record holding the measured value: term1 (scalar, X/Y component) or scale1 (derived angle of a D set)
    uint32_t st[3];      // station indices (unused entries repeat st[0])
    uint32_t edge[3];    // edge words of the pairs (0,1), (0,2), (1,2): slot | bit31 when the FIRST station of the pair is eliminated later
    uint8_t type;        // measurement type letter; 'D' = derived angle of a direction set
    uint8_t nst;         // 1, 2 or 3 stations
    uint8_t comp;        // X / Y cluster rows: Cartesian component 0..2
    uint8_t clustered;   // 1: member of a cluster (assembled by the cluster kernel)
};

// per cluster (D, X, Y)
struct ClusterDesc {
    uint64_t vinv_off;   // n x n row-major symmetric V^-1 in the cluster matrix pool
    uint64_t pair_off;   // ns x ns edge words (row-major; [j1*ns+j2], bit31: j1 eliminated later); diagonal unused
    uint64_t inc_off;    // station -> rows incidence lists: inc_ptr[st0 + j] .. inc_ptr[st0 + j + 1] index into inc[]
    uint32_t row0, n;    // rows [row0, row0 + n) of the row arrays
    uint32_t st0, ns;    // local station list cstn[st0 .. st0 + ns)
    uint8_t type;
    uint8_t pad[7];
};

struct RowOut {
    double l;
    double a[9];
};

// ---- geometry helpers -------------------------------------------------------------------------------
struct LocalFrame {
    double e[3], n[3], u[3];
};
GADJ_HD LocalFrame local_frame(double lat, double lon)
{
    const double sl = sin(lat), cl = cos(lat), so = sin(lon), co = cos(lon);
    LocalFrame f;
    f.e[0] = -so;
    f.e[1] = co;
    f.e[2] = 0.0;
    f.n[0] = -sl * co;
    f.n[1] = -sl * so;
    f.n[2] = cl;
    f.u[0] = cl * co;
    f.u[1] = cl * so;
    f.u[2] = sl;
    return f;
}
GADJ_HD double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// atan_2 (dnatemplatecalcfuncs.hpp:350-362) and Direction (GEO:679-693)
GADJ_HD double ref_atan_2(double x, double y)
{
    const double t = atan(x / y);
    if (y < 0)
        return t + kPi;
    if (x > 0)
        return t;
    return t + kTwoPi;
}
GADJ_HD double direction_en(double e, double n)
{
    double d = fabs(e) < fabs(n) ? ref_atan_2(e, n) : kHalfPi - ref_atan_2(n, e);
    if (d < 0)
        d += kTwoPi;
    return d;
}
// azimuth 1 -> 2 in the frame of station 1, with the local components
GADJ_HD double azimuth(const LocalFrame& f, const double* p1, const double* p2, double* e, double* n)
{
    const double d[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
    *e = dot3(f.e, d);
    *n = dot3(f.n, d);
    return direction_en(*e, *n);
}
// instrument -> target vector (heights along each station's own ellipsoid normal) in the frame of station 1
GADJ_HD void sight_line(const LocalFrame& f1, const double* p1, const double* p2, const double* llh2, double ih, double th,
                        double* e, double* n, double* u)
{
    const double c2 = cos(llh2[0]);
    const double up2[3] = {c2 * cos(llh2[1]), c2 * sin(llh2[1]), sin(llh2[0])};
    const double d[3] = {p2[0] - p1[0] + up2[0] * th - f1.u[0] * ih, p2[1] - p1[1] + up2[1] * th - f1.u[1] * ih,
                         p2[2] - p1[2] + up2[2] * th - f1.u[2] * ih};
    *e = dot3(f1.e, d);
    *n = dot3(f1.n, d);
    *u = dot3(f1.u, d);
}
GADJ_HD double zenith_distance(const LocalFrame& f1, const double* p1, const double* p2, const double* llh2, double ih, double th)
{
    double e, n, u;
    sight_line(f1, p1, p2, llh2, ih, th, &e, &n, &u);
    return atan2(sqrt(e * e + n * n), u);
}
GADJ_HD void nu_rho(const Ellipsoid& el, double lat, double* nu, double* rho)
{
    const double s = sin(lat);
    const double d = sqrt(1.0 - el.e2 * (s * s));
    *nu = el.a / d;
    *rho = el.a * ((1.0 - el.e2) / (d * d * d));
}
// radius of curvature of the ellipsoid in the direction of the chord 1 -> 2 (GEO:993-1008)
GADJ_HD double chord_radius(const Ellipsoid& el, const double* p1, const double* p2, const double* llh1, const double* llh2)
{
    double nu, rho, e, n;
    nu_rho(el, (llh1[0] + llh2[0]) / 2., &nu, &rho);
    const LocalFrame f = local_frame(llh1[0], llh1[1]);
    const double az = azimuth(f, p1, p2, &e, &n);
    const double c = cos(az), s = sin(az);
    return rho * nu / ((nu * c * c) + (rho * s * s));
}
// mean-sea-level arc <-> ellipsoid chord (GEO:1045-1150)
GADJ_HD double msl_arc_to_chord(const Ellipsoid& el, double arc, double lat1, double lat2, double N1, double N2)
{
    double nu, rho;
    nu_rho(el, (lat1 + lat2) / 2., &nu, &rho);
    const double rm = sqrt(nu * rho);
    const double r = rm + (N1 + N2) / 2.;
    const double msl_chord = 2.0 * r * sin(arc / 2.0 / r);
    double c = msl_chord * msl_chord;
    c -= (N2 - N1) * (N2 - N1);
    c /= 1. + N1 / rm;
    c /= 1. + N2 / rm;
    return sqrt(c);
}
GADJ_HD double chord_to_msl_arc(const Ellipsoid& el, double chord, double lat1, double lat2, double N1, double N2)
{
    double nu, rho;
    nu_rho(el, (lat1 + lat2) / 2., &nu, &rho);
    const double rm = sqrt(nu * rho);
    double c = chord * chord;
    c *= 1. + N1 / rm;
    c *= 1. + N2 / rm;
    c += (N2 - N1) * (N2 - N1);
    const double r = rm + (N1 + N2) / 2.;
    return asin(sqrt(c) / 2.0 / r) * 2.0 * r;
}
GADJ_HD double direction_deflection(double az, double zen, double dV, double dM) { return (dM * sin(az) - dV * cos(az)) / tan(zen); }

GADJ_HD double wrap_residual(double mmc)
{   // AddMsrtoMeasMinusComp (ADJ:4718-4735)
    if (mmc < -5.5)
        mmc += kTwoPi;
    else if (mmc > 5.5)
        mmc -= kTwoPi;
    return mmc;
}

// d(azimuth 1->2)/d(p1) in Cartesian components; d/d(p2) is its negative
GADJ_HD void azimuth_gradient(const LocalFrame& f, double e, double n, double* g1)
{
    const double q = 1.0 / (e * e + n * n);
#pragma unroll
    for (int k = 0; k < 3; ++k)
        g1[k] = -(n * f.e[k] - e * f.n[k]) * q;
}

// ---- one design row ---------------------------------------------------------------------------------------
// value: the measured quantity as the record holds it after the first-run reductions (term1, or scale1 for a derived angle).
// est: 3*nstn estimates; llh: 3*nstn current geographic coordinates; geoid: per-station geoid separation (may be null
// when the type does not need it).  Types E and M recompute their reduced value from preAdjMeas on every call (the
// reference does the same, ADJ:5262-5275, ADJ:5414-5421): *reduced receives it (pass nullptr otherwise).
GADJ_HD bool design_row(const RowDesc& d, double value, double term3, double term4, double pre_adj_meas, const double* est,
                        const double* llh, const float* geoid, const Ellipsoid& el, RowOut& o, double* reduced)
{
    const double* p1 = est + 3 * (size_t)d.st[0];
    const double* p2 = est + 3 * (size_t)d.st[1];
    const double* g1 = llh + 3 * (size_t)d.st[0];
    const double* g2 = llh + 3 * (size_t)d.st[1];
#pragma unroll
    for (int k = 0; k < 9; ++k)
        o.a[k] = 0.0;
    switch (d.type) {
    case 'X': {
        o.l = value - (p2[d.comp] - p1[d.comp]);
        o.a[d.comp] = -1.0;
        o.a[3 + d.comp] = 1.0;
        return true;
    }
    case 'Y': {
        o.l = value - p1[d.comp];
        o.a[d.comp] = 1.0;
        return true;
    }
    case 'A':
    case 'D': {
        const double* p3 = est + 3 * (size_t)d.st[2];
        const LocalFrame f = local_frame(g1[0], g1[1]);
        double e12, n12, e13, n13;
        const double d12 = azimuth(f, p1, p2, &e12, &n12);
        double d13 = azimuth(f, p1, p3, &e13, &n13);
        if (d12 > d13)
            d13 += kTwoPi;
        o.l = wrap_residual(value - (d13 - d12));
        double ga[3], gb[3];
        azimuth_gradient(f, e12, n12, ga);   // d az12 / d p1
        azimuth_gradient(f, e13, n13, gb);   // d az13 / d p1
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            o.a[k] = gb[k] - ga[k];
            o.a[3 + k] = ga[k];      // d(-az12)/d p2 = + d az12 / d p1
            o.a[6 + k] = -gb[k];     // d( az13)/d p3 = - d az13 / d p1
        }
        return true;
    }
    case 'B':
    case 'K': {
        const LocalFrame f = local_frame(g1[0], g1[1]);
        double e, n;
        const double az = azimuth(f, p1, p2, &e, &n);
        o.l = wrap_residual(value - az);
        double ga[3];
        azimuth_gradient(f, e, n, ga);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            o.a[k] = ga[k];
            o.a[3 + k] = -ga[k];
        }
        return true;
    }
    case 'C':
    case 'E':
    case 'M': {
        double meas = value;
        if (d.type == 'E') {
            const double r = chord_radius(el, p1, p2, g1, g2);
            meas = 2.0 * r * sin(pre_adj_meas / 2.0 / r);
        } else if (d.type == 'M')
            meas = msl_arc_to_chord(el, pre_adj_meas, g1[0], g2[0], (double)geoid[d.st[0]], (double)geoid[d.st[1]]);
        if (reduced)
            *reduced = meas;
        // chord between the ellipsoid foot points (EllipsoidChordDistance, GEO:958-991)
        const double nu1 = prime_vertical(el, g1[0]), nu2 = prime_vertical(el, g2[0]);
        const double s1 = nu1 / (nu1 + g1[2]), s2 = nu2 / (nu2 + g2[2]);
        const double Zn1 = el.e2 * nu1 * sin(g1[0]), Zn2 = el.e2 * nu2 * sin(g2[0]);
        const double dX = p2[0] * s2 - p1[0] * s1;
        const double dY = p2[1] * s2 - p1[1] * s1;
        const double dZ = ((p2[2] + Zn2) * s2 - Zn2) - ((p1[2] + Zn1) * s1 - Zn1);
        const double comp = sqrt((dX * dX) + (dY * dY) + (dZ * dZ));
        o.l = meas - comp;
        o.a[0] = -dX / comp;
        o.a[1] = -dY / comp;
        o.a[2] = -dZ / comp;
        o.a[3] = -o.a[0];
        o.a[4] = -o.a[1];
        o.a[5] = -o.a[2];
        return true;
    }
    case 'S': {
        // both heights along the normal of station 1 (CartesianElementsFromInstrumentHeight at station 1, ADJ:5450-5457)
        const double cl = cos(g1[0]), sl = sin(g1[0]), co = cos(g1[1]), so = sin(g1[1]);
        const double dX = p2[0] - p1[0] + cl * co * term4 - cl * co * term3;
        const double dY = p2[1] - p1[1] + cl * so * term4 - cl * so * term3;
        const double dZ = p2[2] - p1[2] + sl * term4 - sl * term3;
        const double comp = sqrt(dX * dX + dY * dY + dZ * dZ);
        o.l = value - comp;
        o.a[0] = -dX / comp;
        o.a[1] = -dY / comp;
        o.a[2] = -dZ / comp;
        o.a[3] = -o.a[0];
        o.a[4] = -o.a[1];
        o.a[5] = -o.a[2];
        return true;
    }
    case 'V':
    case 'Z': {
        const LocalFrame f = local_frame(g1[0], g1[1]);
        double e, n, u;
        sight_line(f, p1, p2, g2, term3, term4, &e, &n, &u);
        const double h2 = e * e + n * n, h = sqrt(h2);
        const double comp = d.type == 'V' ? atan2(h, u) : atan2(u, h);
        o.l = value - comp;
        // z = atan2(h, u):  dz/dp2 = (u dh - h du) / (h^2 + u^2), dh = (e E + n N) / h, du = U; the vertical angle is pi/2 - z.
        // (the reference writes the same derivative as cos^2(z) (.. / (u h) + .. h / u^2), ADJ:5578-5590, 5687-5699)
        const double q = 1.0 / (h2 + u * u);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const double dz2 = (u * ((e * f.e[k] + n * f.n[k]) / h) - h * f.u[k]) * q;
            o.a[k] = d.type == 'V' ? -dz2 : dz2;
            o.a[3 + k] = -o.a[k];
        }
        return true;
    }
    case 'L': {
        const double nu1 = prime_vertical(el, g1[0]), nu2 = prime_vertical(el, g2[0]);
        const double Zn1 = el.e2 * nu1 * sin(g1[0]), Zn2 = el.e2 * nu2 * sin(g2[0]);
        const double h1 = sqrt(p1[0] * p1[0] + p1[1] * p1[1] + (p1[2] + Zn1) * (p1[2] + Zn1)) - nu1;
        const double h2 = sqrt(p2[0] * p2[0] + p2[1] * p2[1] + (p2[2] + Zn2) * (p2[2] + Zn2)) - nu2;
        o.l = value - (h2 - h1);
        o.a[0] = -p1[0] / (nu1 + h1);
        o.a[1] = -p1[1] / (nu1 + h1);
        o.a[2] = -(p1[2] + Zn1) / (nu1 + h1);
        o.a[3] = p2[0] / (nu2 + h2);
        o.a[4] = p2[1] / (nu2 + h2);
        o.a[5] = (p2[2] + Zn2) / (nu2 + h2);
        return true;
    }
    case 'H':
    case 'R': {
        const double nu1 = prime_vertical(el, g1[0]);
        const double Zn1 = el.e2 * nu1 * sin(g1[0]);
        const double h1 = sqrt(p1[0] * p1[0] + p1[1] * p1[1] + (p1[2] + Zn1) * (p1[2] + Zn1)) - nu1;
        o.l = value - h1;
        o.a[0] = p1[0] / (nu1 + h1);
        o.a[1] = p1[1] / (nu1 + h1);
        o.a[2] = (p1[2] + Zn1) / (nu1 + h1);
        return true;
    }
    case 'I':
    case 'P': {
        // forward difference of the Lin & Wang latitude with a 1e-4 m increment (PartialD_Latitude, GEO:279-320)
        double g[3];
        cart_to_geo(el, p1[0], p1[1], p1[2], g);
        const double lat = g[0];
        o.l = value - lat;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double c[3] = {p1[0], p1[1], p1[2]};
            c[k] += 1.0e-4;
            cart_to_geo(el, c[0], c[1], c[2], g);
            o.a[k] = (g[0] - lat) / 1.0e-4;
        }
        return true;
    }
    case 'J':
    case 'Q': {
        o.l = value - g1[1];
        const double t = p1[0] * p1[1] / pow(p1[0] * p1[0] + p1[1] * p1[1], 1.5);
        o.a[0] = t * -1. / cos(g1[1]);
        o.a[1] = t / sin(g1[1]);
        o.a[2] = 0.;
        return true;
    }
    default:
        return false;
    }
}

// ---- first-run reductions (host, once per raw measurement file; InitialiseMeasurement ADJ:3913-3935 and the
// "buildnewMatrices && !rebuildingDesign_" branches of each type) ----------------------------------------------
// `m` holds term1/term3/term4/preAdjCorr of the scalar measurement (for a derived angle: the scratch angle record).
inline void first_run_reduce(char type, dna_msr_t& m, const uint32_t st[3], const dna_stn_t* stn, const double* est)
{
    const dna_stn_t& s1 = stn[st[0]];
    const bool defl = std::fabs(s1.verticalDef) > kDeflectionEps || std::fabs(s1.meridianDef) > kDeflectionEps;
    const double* p1 = est + 3 * (size_t)st[0];
    const double* p2 = est + 3 * (size_t)st[1];
    const LocalFrame f = local_frame(s1.currentLatitude, s1.currentLongitude);
    double e, n;
    switch (type) {
    case 'A':
    case 'D': {
        if (!defl) {
            m.preAdjCorr = 0.0;
            break;
        }
        const double* p3 = est + 3 * (size_t)st[2];
        const double llh2[2] = {stn[st[1]].currentLatitude, stn[st[1]].currentLongitude};
        const double llh3[2] = {stn[st[2]].currentLatitude, stn[st[2]].currentLongitude};
        const double d12 = azimuth(f, p1, p2, &e, &n);
        double d13 = azimuth(f, p1, p3, &e, &n);
        if (d12 > d13)
            d13 += kTwoPi;
        const double z12 = zenith_distance(f, p1, p2, llh2, m.term3, m.term4);
        const double z13 = zenith_distance(f, p1, p3, llh3, m.term3, m.term4);
        m.preAdjCorr = direction_deflection(d13, z13, s1.verticalDef, s1.meridianDef) -
                       direction_deflection(d12, z12, s1.verticalDef, s1.meridianDef);
        m.term1 -= m.preAdjCorr;
        break;
    }
    case 'B':
        m.preAdjCorr = 0.0;
        break;
    case 'K': {
        if (!defl) {
            m.preAdjCorr = 0.0;
            break;
        }
        const double llh2[2] = {stn[st[1]].currentLatitude, stn[st[1]].currentLongitude};
        const double az = azimuth(f, p1, p2, &e, &n);
        const double zen = zenith_distance(f, p1, p2, llh2, m.term3, m.term4);
        m.preAdjCorr = s1.verticalDef * tan(s1.currentLatitude) + direction_deflection(az, zen, s1.verticalDef, s1.meridianDef);
        m.term1 -= m.preAdjCorr;
        break;
    }
    case 'C':
        m.preAdjCorr = 0.0;
        break;
    case 'V':
    case 'Z': {
        if (!defl) {
            m.preAdjCorr = 0.0;
            break;
        }
        const double az = azimuth(f, p1, p2, &e, &n);
        m.preAdjCorr = s1.meridianDef * cos(az) + s1.verticalDef * sin(az);
        if (type == 'V')
            m.term1 += m.preAdjCorr;
        else
            m.term1 -= m.preAdjCorr;
        break;
    }
    case 'L': {
        const dna_stn_t& s2 = stn[st[1]];
        if (std::fabs(s1.geoidSep) > 1.0e-4 || std::fabs(s2.geoidSep) > 1.0e-4) {
            m.preAdjCorr = s2.geoidSep - s1.geoidSep;
            m.term1 += m.preAdjCorr;
        }
        break;
    }
    case 'H':
        if (std::fabs(s1.geoidSep) > 1.0e-4) {
            m.preAdjCorr = s1.geoidSep;
            m.term1 += m.preAdjCorr;
        }
        break;
    case 'I':
        if (std::fabs(s1.meridianDef) > kDeflectionEps) {
            m.preAdjCorr = s1.meridianDef;
            m.term1 -= m.preAdjCorr;
        } else
            m.preAdjCorr = 0.0;
        break;
    case 'J':
        if (std::fabs(s1.verticalDef) > kDeflectionEps) {
            m.preAdjCorr = s1.verticalDef / cos(s1.currentLatitude);
            m.term1 -= m.preAdjCorr;
        } else
            m.preAdjCorr = 0.0;
        break;
    default:
        break;   // E, M: reduced on every evaluation; P Q R S: nothing to reduce
    }
}

// ---- statistics of one row (UpdateMsrRecord / UpdateMsrRecordStats, ADJ:8187-8298) --------------------------------
struct RowStats {
    double corr, adj, prec, resid_prec, nstat, pelzer;
    int outlier, reliable;
};
GADJ_HD RowStats row_statistics(double l, double prec, double meas_prec, double critical, double limit)
{
    RowStats s;
    s.corr = -l;
    s.prec = prec;
    double rp = meas_prec - prec;
    if (rp < 0.0)
        rp = fabs(rp);
    s.resid_prec = rp;
    double pel = sqrt(meas_prec) / sqrt(rp);
    if (pel < 0. || pel > 700.)
        pel = 999.99;
    s.nstat = s.corr / sqrt(rp);
    s.outlier = fabs(s.nstat) > critical ? 1 : 0;
    s.reliable = (pel > 0. && pel < limit) ? 1 : 0;   // ComputeGlobalPelzer (ADJ:8302-8427)
    if (!s.reliable)
        pel = 999.99;
    s.pelzer = pel;
    s.adj = 0.0;
    return s;
}


// ---- kernel parameter blocks (device pointers) ----------------------------------------------------------------
struct RowsParams {
    dna_msr_t* msr;              // device records (E / M rows refresh term1 / preAdjCorr; statistics write the result fields)
    const RowDesc* rows;
    uint64_t nrows;
    const double* est;           // 3 x nstn estimated Cartesian coordinates
    const double* llh;           // 3 x nstn current latitude, longitude, height
    const float* geoid;          // nstn geoid separations
    double* row_l;               // nrows: measured - computed
    double* row_a;               // 9 x nrows partials
    double* ndiag;
    double* noff;
    double* w;
    const double* vcv_diag;      // statistics pass: nstn x 9
    const double* vcv_off;       // statistics pass: nedge x 9, N^-1[hi, lo]
    double* sums;                // statistics pass: [0] chi2, [1] pelzer sum, [2] pelzer count, [3] outliers
    double semi_major, inv_flattening, critical;
    int32_t normals;             // 1: add the N contributions of the independent rows, 0: right-hand side only
    int32_t assemble;            // 0: evaluate l and the partials only (statistics pass)
};

struct ClusterParams {
    const ClusterDesc* clusters;
    uint32_t nclusters;
    const double* row_l;
    const double* row_a;
    double* row_t;               // nrows scratch: V^-1 l
    const double* cvinv;         // cluster matrix pool
    const uint32_t* cstn;        // local station lists
    const uint32_t* inc_ptr;     // per local station: range into inc
    const uint32_t* inc;         // (local row << 2) | slot of the station in that row
    const uint32_t* pair_word;   // per cluster ns x ns edge words
    double* ndiag;
    double* noff;
    double* w;
    double* sums;                // statistics: [0] += l^T V^-1 l of the X / Y clusters
    int32_t normals;
};

#if defined(__CUDA_ARCH__)
#define GADJ_ACC(ptr, v) atomicAdd((ptr), (v))
#else
#define GADJ_ACC(ptr, v) (*(ptr) += (v))
#endif

// block N[sa, sb] += B (3x3, rows belong to sa) into the stored lower block of the pair; word bit31: sa is the row station
GADJ_HD void add_pair_block(double* noff, uint32_t word, const double* B)
{
    double* o = noff + 9 * (size_t)(word & 0x7FFFFFFFu);
    if (word & 0x80000000u) {
#pragma unroll
        for (int k = 0; k < 9; ++k)
            GADJ_ACC(o + k, B[k]);
    } else {
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                GADJ_ACC(o + 3 * b + a, B[3 * a + b]);
    }
}
// Q[sa, sb] (3x3, rows belong to sa) from the stored block of the pair
GADJ_HD double pair_block_at(const double* qoff, uint32_t word, int a, int b)
{
    const double* q = qoff + 9 * (size_t)(word & 0x7FFFFFFFu);
    return (word & 0x80000000u) ? q[3 * a + b] : q[3 * b + a];
}
GADJ_HD uint32_t pair_slot_of(int i, int j) { return i == 0 ? (j == 1 ? 0u : 1u) : 2u; }   // (0,1) (0,2) (1,2)

GADJ_HD double row_value(const RowDesc& d, const dna_msr_t& m) { return d.type == 'D' ? m.scale1 : m.term1; }
GADJ_HD double row_variance(const RowDesc& d, const dna_msr_t& m)
{
    if (d.type == 'D')
        return m.scale2;
    if (d.type == 'X' || d.type == 'Y')
        return d.comp == 0 ? m.term2 : (d.comp == 1 ? m.term3 : m.term4);
    return m.term2;
}

// one row: evaluate, store, and (independent rows) accumulate p a^T a and p a^T l
GADJ_HD void row_body(const RowsParams& p, uint64_t i, const Ellipsoid& el)
{
    const RowDesc d = p.rows[i];
    dna_msr_t* m = p.msr + d.rec;
    RowOut o;
    double reduced = 0.0;
    if (!design_row(d, row_value(d, *m), m->term3, m->term4, m->preAdjMeas, p.est, p.llh, p.geoid, el, o, &reduced))
        return;
    if (d.type == 'E' || d.type == 'M') {
        m->term1 = reduced;
        m->preAdjCorr = reduced - m->preAdjMeas;
    }
    p.row_l[i] = o.l;
#pragma unroll
    for (int k = 0; k < 9; ++k)
        p.row_a[9 * i + k] = o.a[k];
    if (!p.assemble || d.clustered)
        return;
    const double wt = 1.0 / m->term2;   // UpdateAtVinv (ADJ:1285-1320)
    for (int s = 0; s < d.nst; ++s) {
        double* ws = p.w + 3 * (size_t)d.st[s];
#pragma unroll
        for (int k = 0; k < 3; ++k)
            GADJ_ACC(ws + k, wt * o.a[3 * s + k] * o.l);
    }
    if (!p.normals)
        return;
    for (int s = 0; s < d.nst; ++s) {
        double* ds = p.ndiag + 9 * (size_t)d.st[s];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                GADJ_ACC(ds + 3 * a + b, wt * o.a[3 * s + a] * o.a[3 * s + b]);
        for (int t = s + 1; t < d.nst; ++t) {
            double B[9];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    B[3 * a + b] = wt * o.a[3 * s + a] * o.a[3 * t + b];
            add_pair_block(p.noff, d.edge[pair_slot_of(s, t)], B);
        }
    }
}

// statistics of one row; acc[0..3] += chi2, pelzer sum, pelzer count, outliers
GADJ_HD void row_stats_body(const RowsParams& p, uint64_t i, const Ellipsoid& el, double* acc)
{
    const RowDesc d = p.rows[i];
    dna_msr_t* m = p.msr + d.rec;
    const double l = p.row_l[i];
    const double* a = p.row_a + 9 * i;
    double prec;
    if (d.type == 'X') {
        // Precision_Adjusted_GNSS_bsl (MFN:255-297), diagonal term of component comp
        const int c = d.comp;
        const double q21 = (p.vcv_off + 9 * (size_t)(d.edge[0] & 0x7FFFFFFFu))[4 * c];
        const double t0 = -p.vcv_diag[9 * (size_t)d.st[0] + 4 * c] + q21;
        const double tt = -q21 + p.vcv_diag[9 * (size_t)d.st[1] + 4 * c];
        prec = tt - t0;
    } else if (d.type == 'Y') {
        prec = p.vcv_diag[9 * (size_t)d.st[0] + 4 * d.comp];   // ComputePrecisionAdjMsrs_Y (ADJ:8035-8060)
    } else {
        // a Q a^T over the row's stations (ComputePrecisionAdjMsrs_A / _BCEKLMSVZ / _HIJPQR, ADJ:7880-8003)
        prec = 0.0;
        for (int s = 0; s < d.nst; ++s) {
            const double* Q = p.vcv_diag + 9 * (size_t)d.st[s];
            for (int x = 0; x < 3; ++x)
                for (int y = 0; y < 3; ++y)
                    prec += a[3 * s + x] * Q[3 * x + y] * a[3 * s + y];
            for (int t = s + 1; t < d.nst; ++t) {
                const uint32_t wd = d.edge[pair_slot_of(s, t)];
                for (int x = 0; x < 3; ++x)
                    for (int y = 0; y < 3; ++y)
                        prec += 2.0 * a[3 * s + x] * pair_block_at(p.vcv_off, wd, x, y) * a[3 * t + y];
            }
        }
    }
    const double var = row_variance(d, *m);
    RowStats s = row_statistics(l, prec, var, p.critical, 700.0);
    double adj = row_value(d, *m) + s.corr;
    switch (d.type) {   // ADJ:8194-8271
    case 'D':
        if (adj > kTwoPi)
            adj -= kTwoPi;
        adj += m->preAdjCorr;
        break;
    case 'E': {
        const double r = chord_radius(el, p.est + 3 * (size_t)d.st[0], p.est + 3 * (size_t)d.st[1], p.llh + 3 * (size_t)d.st[0],
                                      p.llh + 3 * (size_t)d.st[1]);
        adj = asin(adj / 2.0 / r) * 2.0 * r;
        break;
    }
    case 'M':
        adj = chord_to_msl_arc(el, adj, p.llh[3 * (size_t)d.st[0]], p.llh[3 * (size_t)d.st[1]], (double)p.geoid[d.st[0]],
                               (double)p.geoid[d.st[1]]);
        break;
    case 'H':
    case 'L':
    case 'V':
        adj -= m->preAdjCorr;
        break;
    case 'A':
    case 'I':
    case 'J':
    case 'K':
    case 'Z':
        adj += m->preAdjCorr;
        break;
    default:
        break;
    }
    m->measCorr = s.corr;
    m->measAdj = adj;
    m->measAdjPrec = s.prec;
    m->residualPrec = s.resid_prec;
    m->NStat = s.nstat;
    m->PelzerRel = s.pelzer;
    if (d.type != 'X' && d.type != 'Y')
        acc[0] += l * l / var;   // ComputeChiSquare_ABCEHIJKLMPQRSVZ / _D (ADJ:8430-8469); X / Y: cluster_chi_body
    if (s.reliable) {
        acc[1] += s.pelzer * s.pelzer - 1.;
        acc[2] += 1.0;
    }
    acc[3] += s.outlier;
}

// ---- clusters ------------------------------------------------------------------------------------------------
// t[r] = sum_r' V^-1[r][r'] l[r']
GADJ_HD void cluster_t_body(const ClusterParams& p, const ClusterDesc& c, uint32_t r)
{
    const double* V = p.cvinv + c.vinv_off + (size_t)r * c.n;
    const double* l = p.row_l + c.row0;
    double t = 0.0;
    for (uint32_t q = 0; q < c.n; ++q)
        t += V[q] * l[q];
    p.row_t[c.row0 + r] = t;
}
// w[station j] += sum over the rows that contain j of a_r[j] t[r]
GADJ_HD void cluster_rhs_body(const ClusterParams& p, const ClusterDesc& c, uint32_t j)
{
    const uint32_t b = p.inc_ptr[c.st0 + j], e = p.inc_ptr[c.st0 + j + 1];
    double acc[3] = {0.0, 0.0, 0.0};
    for (uint32_t x = b; x < e; ++x) {
        const uint32_t r = p.inc[x] >> 2, k = p.inc[x] & 3u;
        const double* a = p.row_a + 9 * (size_t)(c.row0 + r) + 3 * k;
        const double t = p.row_t[c.row0 + r];
        acc[0] += a[0] * t;
        acc[1] += a[1] * t;
        acc[2] += a[2] * t;
    }
    double* ws = p.w + 3 * (size_t)p.cstn[c.st0 + j];
    GADJ_ACC(ws + 0, acc[0]);
    GADJ_ACC(ws + 1, acc[1]);
    GADJ_ACC(ws + 2, acc[2]);
}
// N[station j1, station j2] += sum_{r ∋ j1} sum_{r' ∋ j2} V^-1[r][r'] a_r[j1]^T a_r'[j2]   (pair index -> j1 >= j2)
GADJ_HD void cluster_pair_body(const ClusterParams& p, const ClusterDesc& c, uint64_t pair)
{
    // pair = j1 (j1 + 1) / 2 + j2
    uint32_t j1 = (uint32_t)((sqrt(8.0 * (double)pair + 1.0) - 1.0) * 0.5);
    while ((uint64_t)j1 * (j1 + 1) / 2 > pair)
        --j1;
    while ((uint64_t)(j1 + 1) * (j1 + 2) / 2 <= pair)
        ++j1;
    const uint32_t j2 = (uint32_t)(pair - (uint64_t)j1 * (j1 + 1) / 2);
    const uint32_t b1 = p.inc_ptr[c.st0 + j1], e1 = p.inc_ptr[c.st0 + j1 + 1];
    const uint32_t b2 = p.inc_ptr[c.st0 + j2], e2 = p.inc_ptr[c.st0 + j2 + 1];
    double B[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t x = b1; x < e1; ++x) {
        const uint32_t r = p.inc[x] >> 2, k1 = p.inc[x] & 3u;
        const double* a1 = p.row_a + 9 * (size_t)(c.row0 + r) + 3 * k1;
        const double* V = p.cvinv + c.vinv_off + (size_t)r * c.n;
        double u[3] = {0.0, 0.0, 0.0};   // sum_r' V[r][r'] a_r'[j2]
        for (uint32_t y = b2; y < e2; ++y) {
            const uint32_t r2 = p.inc[y] >> 2, k2 = p.inc[y] & 3u;
            const double* a2 = p.row_a + 9 * (size_t)(c.row0 + r2) + 3 * k2;
            const double v = V[r2];
            u[0] += v * a2[0];
            u[1] += v * a2[1];
            u[2] += v * a2[2];
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                B[3 * a + b] += a1[a] * u[b];
    }
    if (j1 == j2) {
        double* ds = p.ndiag + 9 * (size_t)p.cstn[c.st0 + j1];
#pragma unroll
        for (int k = 0; k < 9; ++k)
            GADJ_ACC(ds + k, B[k]);
    } else
        add_pair_block(p.noff, p.pair_word[c.pair_off + (size_t)j1 * c.ns + j2], B);
}

}  // namespace gadj
