// oracle/oracle_types.h — TEST INFRASTRUCTURE (part of the parity checker; included by oracle.cpp only).
//
// Restatement of the reference's per-measurement-type arithmetic for the terrestrial rows:
// computed value, measured-minus-computed, partial derivatives and first-run reductions of
//   A  horizontal angle          UpdateDesignNormalMeasMatrices_A   ADJ:4754-4910
//   B/K geodetic/astro azimuth   UpdateDesignNormalMeasMatrices_BK  ADJ:4913-5014
//   C/E/M chord / ellipsoid arc / MSL arc   _C ADJ:5017-5033, _E ADJ:5242-5281, _M ADJ:5398-5428, _CEM ADJ:5036-5079
//   V/Z zenith distance / vertical angle     ADJ:5504-5601, ADJ:5613-5710
//   H/R orthometric / ellipsoidal height     ADJ:5969-6053
//   I/P astronomic / geodetic latitude       ADJ:5786-5812, 5846-5914
//   J/Q astronomic / geodetic longitude      ADJ:5816-5843, 5917-5966
//   S/L slope distance / level difference    ADJ:5437-5493, ADJ:5717-5784
// Geodesy helpers follow include/functions/dnatemplategeodesyfuncs.hpp (GEO) line by line.
// This file lives inside oracle.cpp's anonymous namespace (it uses Ellipsoid, prime_vertical, PI ...).

const double HALF_PI = PI / 2.0;
const double PRECISION_1E4 = 1.0e-4;
const double SEC_TO_RAD = PI / 180.0 / 3600.0;
const double E4_SEC_DEFLECTION = 0.0001 * SEC_TO_RAD;   // dnaconsts.hpp:110

// one design row: residual, partials w.r.t. up to three stations
struct Row {
    uint32_t st[3] = {0, 0, 0};
    int nst = 0;
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
};

// atan_2 (dnatemplatecalcfuncs.hpp:350-362)
inline double atan_2(double x, double y)
{
    double theta = std::atan(x / y);
    if (y < 0)
        return theta + PI;
    if (x > 0)
        return theta;
    return theta + TWO_PI;
}

// ComputeLocalElements3D (GEO:628-652)
inline void local_elements(double X1, double Y1, double Z1, double X2, double Y2, double Z2, double lat, double lon, double* e,
                           double* n, double* up)
{
    double dX = X2 - X1, dY = Y2 - Y1, dZ = Z2 - Z1;
    double sin_lat = std::sin(lat), cos_lat = std::cos(lat), sin_long = std::sin(lon), cos_long = std::cos(lon);
    *e = -sin_long * dX + cos_long * dY;
    *n = -sin_lat * cos_long * dX - sin_lat * sin_long * dY + cos_lat * dZ;
    if (up)
        *up = cos_lat * cos_long * dX + cos_lat * sin_long * dY + sin_lat * dZ;
}

// Direction (GEO:679-711)
inline double direction_en(double e, double n)
{
    double d;
    if (std::fabs(e) < std::fabs(n))
        d = atan_2(e, n);
    else
        d = HALF_PI - atan_2(n, e);
    if (d < 0)
        d += TWO_PI;
    return d;
}
inline double direction(const double* p1, const double* p2, double lat, double lon, double* e, double* n)
{
    local_elements(p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], lat, lon, e, n, nullptr);
    return direction_en(*e, *n);
}

// ZenithDistance / VerticalAngle (GEO:786-887): heights rotated at their own stations
inline void local_elements_heights(const double* p1, const double* p2, double lat1, double lon1, double lat2, double lon2,
                                   double ih, double th, double* e, double* n, double* up)
{
    double sin_lat1 = std::sin(lat1), cos_lat1 = std::cos(lat1), sin_long1 = std::sin(lon1), cos_long1 = std::cos(lon1);
    double dXih = std::cos(lat1) * std::cos(lon1) * ih, dYih = std::cos(lat1) * std::sin(lon1) * ih, dZih = std::sin(lat1) * ih;
    double dXth = std::cos(lat2) * std::cos(lon2) * th, dYth = std::cos(lat2) * std::sin(lon2) * th, dZth = std::sin(lat2) * th;
    double dX = p2[0] - p1[0] + dXth - dXih;
    double dY = p2[1] - p1[1] + dYth - dYih;
    double dZ = p2[2] - p1[2] + dZth - dZih;
    *e = -sin_long1 * dX + cos_long1 * dY;
    *n = -sin_lat1 * cos_long1 * dX - sin_lat1 * sin_long1 * dY + cos_lat1 * dZ;
    *up = cos_lat1 * cos_long1 * dX + cos_lat1 * sin_long1 * dY + sin_lat1 * dZ;
}
inline double zenith_distance(const double* p1, const double* p2, double lat1, double lon1, double lat2, double lon2, double ih,
                              double th, double* e, double* n, double* up)
{
    local_elements_heights(p1, p2, lat1, lon1, lat2, lon2, ih, th, e, n, up);
    return std::atan2(std::sqrt((*e) * (*e) + (*n) * (*n)), *up);
}
inline double vertical_angle(const double* p1, const double* p2, double lat1, double lon1, double lat2, double lon2, double ih,
                             double th, double* e, double* n, double* up)
{
    local_elements_heights(p1, p2, lat1, lon1, lat2, lon2, ih, th, e, n, up);
    return std::atan2(*up, std::sqrt((*e) * (*e) + (*n) * (*n)));
}

// primeVerticalandMeridian_ (dnadatumprojectionparam.hpp:79-86)
inline void nu_rho(const Ellipsoid& el, double lat, double* nu, double* rho)
{
    double dDel = std::sqrt(1.0 - el.e2 * (std::sin(lat) * std::sin(lat)));
    *nu = el.a / dDel;
    *rho = el.a * ((1.0 - el.e2) / (dDel * dDel * dDel));
}
inline double average2(double a, double b)
{
    double t = a + b;
    return t / 2.;
}

// EllipsoidChordDistance (GEO:958-991)
inline double ellipsoid_chord(const Ellipsoid& el, const double* p1, const double* p2, double lat1, double lat2, double h1, double h2,
                              double* dX, double* dY, double* dZ)
{
    double nu1 = prime_vertical(el, lat1), nu2 = prime_vertical(el, lat2);
    double scale1 = nu1 / (nu1 + h1), scale2 = nu2 / (nu2 + h2);
    double Zn1 = el.e2 * nu1 * std::sin(lat1), Zn2 = el.e2 * nu2 * std::sin(lat2);
    double x1 = p1[0] * scale1, y1 = p1[1] * scale1, z1 = (p1[2] + Zn1) * scale1 - Zn1;
    double x2 = p2[0] * scale2, y2 = p2[1] * scale2, z2 = (p2[2] + Zn2) * scale2 - Zn2;
    *dX = x2 - x1;
    *dY = y2 - y1;
    *dZ = z2 - z1;
    return std::sqrt(((*dX) * (*dX)) + ((*dY) * (*dY)) + ((*dZ) * (*dZ)));
}

// RadiusCurvatureInChordDirection (GEO:993-1008)
inline double radius_in_chord_direction(const Ellipsoid& el, const double* p1, const double* p2, double lat1, double lon1, double lat2)
{
    double nu, rho, e, n;
    nu_rho(el, average2(lat1, lat2), &nu, &rho);
    double d = direction(p1, p2, lat1, lon1, &e, &n);
    double cos_dir = std::cos(d), sin_dir = std::sin(d);
    return rho * nu / ((nu * cos_dir * cos_dir) + (rho * sin_dir * sin_dir));
}
// EllipsoidArctoEllipsoidChord / EllipsoidChordtoEllipsoidArc (GEO:1010-1031)
inline double ell_arc_to_chord(const Ellipsoid& el, double arc, const double* p1, const double* p2, double lat1, double lon1, double lat2)
{
    double r = radius_in_chord_direction(el, p1, p2, lat1, lon1, lat2);
    return 2.0 * r * std::sin(arc / 2.0 / r);
}
inline double ell_chord_to_arc(const Ellipsoid& el, double chord, const double* p1, const double* p2, double lat1, double lon1,
                               double lat2)
{
    double r = radius_in_chord_direction(el, p1, p2, lat1, lon1, lat2);
    return std::asin(chord / 2.0 / r) * 2.0 * r;
}
// MSLArctoEllipsoidChord (GEO:1070-1115) and EllipsoidChordtoMSLArc (GEO:1117-1150)
inline double msl_arc_to_ell_chord(const Ellipsoid& el, double arc, double lat1, double lat2, double N1, double N2)
{
    double nu, rho;
    nu_rho(el, average2(lat1, lat2), &nu, &rho);
    double r = std::sqrt(nu * rho) + average2(N1, N2);
    double msl_chord = 2.0 * r * std::sin(arc / 2.0 / r);
    double c = msl_chord * msl_chord;
    c -= std::pow(N2 - N1, 2);
    double meanLat = average2(lat1, lat2);
    double nu2, rho2;
    nu_rho(el, meanLat, &nu2, &rho2);
    c /= 1. + N1 / std::sqrt(nu2 * rho2);
    c /= 1. + N2 / std::sqrt(nu2 * rho2);
    return std::sqrt(c);
}
inline double ell_chord_to_msl_arc(const Ellipsoid& el, double chord, double lat1, double lat2, double N1, double N2)
{
    double c = chord * chord;
    double meanLat = average2(lat1, lat2);
    double nu, rho;
    nu_rho(el, meanLat, &nu, &rho);
    c *= 1. + N1 / std::sqrt(nu * rho);
    c *= 1. + N2 / std::sqrt(nu * rho);
    c += std::pow(N2 - N1, 2);
    double msl_chord = std::sqrt(c);
    double r = std::sqrt(nu * rho) + average2(N1, N2);
    return std::asin(msl_chord / 2.0 / r) * 2.0 * r;
}

// deflection-of-the-vertical corrections (GEO:1170-1208)
inline double laplace_correction(double az, double zen, double dV, double dM, double lat)
{
    return dV * std::tan(lat) + ((dM * std::sin(az) - dV * std::cos(az)) / std::tan(zen));
}
inline double zenith_deflection_correction(double az, double dV, double dM) { return dM * std::cos(az) + dV * std::sin(az); }
inline double direction_deflection_correction(double az, double zen, double dV, double dM)
{
    return (dM * std::sin(az) - dV * std::cos(az)) / std::tan(zen);
}

// CartToLat (GEO:228-277)
inline double cart_to_lat(const Ellipsoid& el, double x, double y, double z)
{
    double lat, lon, h;
    cart_to_geo(el, x, y, z, &lat, &lon, &h);   // identical Newton iteration; latitude is its first output
    return lat;
}

// angular residual wrap (AddMsrtoMeasMinusComp, ADJ:4718-4747)
inline double meas_minus_comp(char type, double term1, double comp)
{
    double mmc = term1 - comp;
    switch (type) {
    case 'A':
    case 'B':
    case 'D':
    case 'K':
        if (mmc < -5.5)
            mmc += TWO_PI;
        else if (mmc > 5.5)
            mmc -= TWO_PI;
    }
    return mmc;
}

// Horizontal angle 1 -> 2 -> 3 (UpdateDesignNormalMeasMatrices_A, ADJ:4754-4910).
// `m` is the record that carries term1/term3/term4 (for a D set: the derived-angle scratch record).
// first_run: apply the Laplace correction to m->term1 (ADJ:4788-4846).
inline double angle_row(const Ellipsoid& el, dna_msr_t* m, const dna_stn_t* stn, const double* est, uint32_t s1, uint32_t s2,
                        uint32_t s3, bool first_run, Row& r)
{
    (void)el;
    const double* p1 = est + 3 * (size_t)s1;
    const double* p2 = est + 3 * (size_t)s2;
    const double* p3 = est + 3 * (size_t)s3;
    const dna_stn_t& st1 = stn[s1];
    double e12, n12, e13, n13;
    double d12 = direction(p1, p2, st1.currentLatitude, st1.currentLongitude, &e12, &n12);
    double d13 = direction(p1, p3, st1.currentLatitude, st1.currentLongitude, &e13, &n13);
    if (d12 > d13)
        d13 += TWO_PI;
    double comp = d13 - d12;
    if (first_run) {
        if (std::fabs(st1.verticalDef) > E4_SEC_DEFLECTION || std::fabs(st1.meridianDef) > E4_SEC_DEFLECTION) {
            double e, n, up;
            double z12 = zenith_distance(p1, p2, st1.currentLatitude, st1.currentLongitude, stn[s2].currentLatitude,
                                         stn[s2].currentLongitude, m->term3, m->term4, &e, &n, &up);
            double z13 = zenith_distance(p1, p3, st1.currentLatitude, st1.currentLongitude, stn[s3].currentLatitude,
                                         stn[s3].currentLongitude, m->term3, m->term4, &e, &n, &up);
            m->preAdjCorr = direction_deflection_correction(d13, z13, st1.verticalDef, st1.meridianDef) -
                            direction_deflection_correction(d12, z12, st1.verticalDef, st1.meridianDef);
            m->term1 -= m->preAdjCorr;
        } else
            m->preAdjCorr = 0.0;
    }
    double l = meas_minus_comp(m->measType, m->term1, comp);
    double cos_lat = std::cos(st1.currentLatitude), sin_lat = std::sin(st1.currentLatitude);
    double cos_long = std::cos(st1.currentLongitude), sin_long = std::sin(st1.currentLongitude);
    double sinlat_coslong = sin_lat * cos_long, sinlat_sinlong = sin_lat * sin_long;
    double c12 = std::cos(d12) * std::cos(d12) / (n12 * n12);
    double c13 = std::cos(d13) * std::cos(d13) / (n13 * n13);
    r.nst = 3;
    r.st[0] = s1;
    r.st[1] = s2;
    r.st[2] = s3;
    r.a[0] = c13 * (n13 * sin_long - e13 * sinlat_coslong) - c12 * (n12 * sin_long - e12 * sinlat_coslong);
    r.a[1] = c13 * (-n13 * cos_long - e13 * sinlat_sinlong) - c12 * (-n12 * cos_long - e12 * sinlat_sinlong);
    r.a[2] = c13 * e13 * cos_lat - c12 * e12 * cos_lat;
    r.a[3] = c12 * (n12 * sin_long - e12 * sinlat_coslong);
    r.a[4] = c12 * (-n12 * cos_long - e12 * sinlat_sinlong);
    r.a[5] = c12 * e12 * cos_lat;
    r.a[6] = -c13 * (n13 * sin_long - e13 * sinlat_coslong);
    r.a[7] = -c13 * (-n13 * cos_long - e13 * sinlat_sinlong);
    r.a[8] = -c13 * e13 * cos_lat;
    return l;
}

inline void two_station(Row& r, uint32_t s1, uint32_t s2, double dx, double dy, double dz)
{   // AddMsrtoDesign_BCEKMSVZ (ADJ:4709-4716)
    r.nst = 2;
    r.st[0] = s1;
    r.st[1] = s2;
    r.a[0] = dx;
    r.a[1] = dy;
    r.a[2] = dz;
    r.a[3] = -dx;
    r.a[4] = -dy;
    r.a[5] = -dz;
}

// One scalar measurement: first-run reductions (first_run), residual and partials.  Returns false for an unknown type.
inline bool scalar_row_oracle(const Ellipsoid& el, dna_msr_t* m, const dna_stn_t* stn, const double* est, bool first_run, Row& r,
                              double* l_out)
{
    const uint32_t s1 = m->station1, s2 = m->station2;
    const double* p1 = est + 3 * (size_t)s1;
    const double* p2 = est + 3 * (size_t)s2;
    const dna_stn_t& st1 = stn[s1];
    const bool defl = std::fabs(st1.verticalDef) > E4_SEC_DEFLECTION || std::fabs(st1.meridianDef) > E4_SEC_DEFLECTION;
    if (first_run)
        m->preAdjMeas = m->term1;   // InitialiseMeasurement (ADJ:3913-3935), first adjustment of raw data
    switch (m->measType) {
    case 'A':
        *l_out = angle_row(el, m, stn, est, s1, s2, m->station3, first_run, r);
        return true;
    case 'B':
    case 'K': {
        const dna_stn_t& st2 = stn[s2];
        double e12, n12;
        double comp = direction(p1, p2, st1.currentLatitude, st1.currentLongitude, &e12, &n12);
        if (first_run) {
            if (m->measType == 'K' && defl) {
                double e, n, up;
                double zen = zenith_distance(p1, p2, st1.currentLatitude, st1.currentLongitude, st2.currentLatitude,
                                             st2.currentLongitude, m->term3, m->term4, &e, &n, &up);
                m->preAdjCorr = laplace_correction(comp, zen, st1.verticalDef, st1.meridianDef, st1.currentLatitude);
                m->term1 -= m->preAdjCorr;
            } else
                m->preAdjCorr = 0.0;
        }
        *l_out = meas_minus_comp(m->measType, m->term1, comp);
        double cos_lat = std::cos(st1.currentLatitude), sin_lat = std::sin(st1.currentLatitude);
        double cos_long = std::cos(st1.currentLongitude), sin_long = std::sin(st1.currentLongitude);
        double c12 = std::cos(comp) * std::cos(comp) / (n12 * n12);
        two_station(r, s1, s2, c12 * (n12 * sin_long - e12 * (sin_lat * cos_long)), c12 * (-n12 * cos_long - e12 * (sin_lat * sin_long)),
                    c12 * e12 * cos_lat);
        return true;
    }
    case 'C':
    case 'E':
    case 'M': {
        const dna_stn_t& st2 = stn[s2];
        if (m->measType == 'C') {
            if (first_run)
                m->preAdjCorr = 0.;
        } else if (m->measType == 'E') {
            m->term1 = ell_arc_to_chord(el, m->preAdjMeas, p1, p2, st1.currentLatitude, st1.currentLongitude, st2.currentLatitude);
            m->preAdjCorr = m->term1 - m->preAdjMeas;
        } else {
            m->term1 = msl_arc_to_ell_chord(el, m->preAdjMeas, st1.currentLatitude, st2.currentLatitude, st1.geoidSep, st2.geoidSep);
            m->preAdjCorr = m->term1 - m->preAdjMeas;
        }
        double dX, dY, dZ;
        double comp = ellipsoid_chord(el, p1, p2, st1.currentLatitude, st2.currentLatitude, st1.currentHeight, st2.currentHeight, &dX,
                                      &dY, &dZ);
        *l_out = meas_minus_comp(m->measType, m->term1, comp);
        two_station(r, s1, s2, -dX / comp, -dY / comp, -dZ / comp);
        return true;
    }
    case 'V':
    case 'Z': {
        const dna_stn_t& st2 = stn[s2];
        if (first_run) {
            if (defl) {
                double e, n;
                double az = direction(p1, p2, st1.currentLatitude, st1.currentLongitude, &e, &n);
                m->preAdjCorr = zenith_deflection_correction(az, st1.verticalDef, st1.meridianDef);
                if (m->measType == 'V')
                    m->term1 += m->preAdjCorr;
                else
                    m->term1 -= m->preAdjCorr;
            } else
                m->preAdjCorr = 0.0;
        }
        double e, n, up;
        double cos_lat = std::cos(st1.currentLatitude), sin_lat = std::sin(st1.currentLatitude);
        double cos_long = std::cos(st1.currentLongitude), sin_long = std::sin(st1.currentLongitude);
        if (m->measType == 'V') {
            double comp = zenith_distance(p1, p2, st1.currentLatitude, st1.currentLongitude, st2.currentLatitude, st2.currentLongitude,
                                          m->term3, m->term4, &e, &n, &up);
            *l_out = meas_minus_comp('V', m->term1, comp);
            double e2n2 = e * e + n * n, sqrt_e2n2 = std::sqrt(e2n2);
            double se2n2_up2 = sqrt_e2n2 / (up * up), up_se2n2 = up * sqrt_e2n2;
            double cos2v = std::cos(comp) * std::cos(comp);
            two_station(r, s1, s2, cos2v * (((e * sin_long + n * sin_lat * cos_long) / up_se2n2) + cos_lat * cos_long * se2n2_up2),
                        cos2v * (((-e * cos_long + n * sin_lat * sin_long) / up_se2n2) + cos_lat * sin_long * se2n2_up2),
                        cos2v * ((-n * cos_lat / up_se2n2) + sin_lat * se2n2_up2));
        } else {
            double comp = vertical_angle(p1, p2, st1.currentLatitude, st1.currentLongitude, st2.currentLatitude, st2.currentLongitude,
                                         m->term3, m->term4, &e, &n, &up);
            *l_out = meas_minus_comp('Z', m->term1, comp);
            double e2n2 = e * e + n * n, sqrt_e2n2 = std::sqrt(e2n2);
            double se2n2_d_e2n2 = sqrt_e2n2 / e2n2, up_d = up / (sqrt_e2n2 * e2n2);
            double cos2v = std::cos(comp) * std::cos(comp);
            two_station(r, s1, s2, cos2v * ((-cos_lat * cos_long * se2n2_d_e2n2) - ((e * sin_long + n * sin_lat * cos_long) * up_d)),
                        cos2v * ((-cos_lat * sin_long * se2n2_d_e2n2) + ((e * cos_long - n * sin_lat * sin_long) * up_d)),
                        cos2v * ((-sin_lat * se2n2_d_e2n2) + (n * cos_lat * up_d)));
        }
        return true;
    }
    case 'S': {
        double cl = std::cos(st1.currentLatitude), sl = std::sin(st1.currentLatitude);
        double co = std::cos(st1.currentLongitude), so = std::sin(st1.currentLongitude);
        // CartesianElementsFromInstrumentHeight (GEO:763-771): both heights are rotated at station 1 (ADJ:5450-5457)
        double dXih = cl * co * m->term3, dYih = cl * so * m->term3, dZih = sl * m->term3;
        double dXth = cl * co * m->term4, dYth = cl * so * m->term4, dZth = sl * m->term4;
        double dX = p2[0] - p1[0] + dXth - dXih;
        double dY = p2[1] - p1[1] + dYth - dYih;
        double dZ = p2[2] - p1[2] + dZth - dZih;
        double comp = std::sqrt(dX * dX + dY * dY + dZ * dZ);
        *l_out = m->term1 - comp;
        two_station(r, s1, s2, -dX / comp, -dY / comp, -dZ / comp);
        return true;
    }
    case 'L': {
        const dna_stn_t& st2 = stn[s2];
        double nu1 = prime_vertical(el, st1.currentLatitude), nu2 = prime_vertical(el, st2.currentLatitude);
        double Zn1 = el.e2 * nu1 * std::sin(st1.currentLatitude), Zn2 = el.e2 * nu2 * std::sin(st2.currentLatitude);
        double h1 = std::sqrt(p1[0] * p1[0] + p1[1] * p1[1] + std::pow(p1[2] + Zn1, 2)) - nu1;
        double h2 = std::sqrt(p2[0] * p2[0] + p2[1] * p2[1] + std::pow(p2[2] + Zn2, 2)) - nu2;
        if (first_run)
            if (std::fabs(st1.geoidSep) > PRECISION_1E4 || std::fabs(st2.geoidSep) > PRECISION_1E4) {
                m->preAdjCorr = st2.geoidSep - st1.geoidSep;
                m->term1 += m->preAdjCorr;
            }
        *l_out = m->term1 - (h2 - h1);
        r.nst = 2;
        r.st[0] = s1;
        r.st[1] = s2;
        r.a[0] = -p1[0] / (nu1 + h1);
        r.a[1] = -p1[1] / (nu1 + h1);
        r.a[2] = -(p1[2] + Zn1) / (nu1 + h1);
        r.a[3] = p2[0] / (nu2 + h2);
        r.a[4] = p2[1] / (nu2 + h2);
        r.a[5] = (p2[2] + Zn2) / (nu2 + h2);
        return true;
    }
    case 'H':
    case 'R': {
        if (first_run && m->measType == 'H')
            if (std::fabs(st1.geoidSep) > PRECISION_1E4) {
                m->preAdjCorr = st1.geoidSep;
                m->term1 += m->preAdjCorr;
            }
        double nu1 = prime_vertical(el, st1.currentLatitude);
        double Zn1 = el.e2 * nu1 * std::sin(st1.currentLatitude);
        double comp = std::sqrt(p1[0] * p1[0] + p1[1] * p1[1] + std::pow(p1[2] + Zn1, 2)) - nu1;
        *l_out = m->term1 - comp;
        r.nst = 1;
        r.st[0] = s1;
        r.a[0] = p1[0] / (nu1 + comp);
        r.a[1] = p1[1] / (nu1 + comp);
        r.a[2] = (p1[2] + Zn1) / (nu1 + comp);
        return true;
    }
    case 'I':
    case 'P': {
        if (first_run && m->measType == 'I') {
            if (std::fabs(st1.meridianDef) > E4_SEC_DEFLECTION) {
                m->preAdjCorr = st1.meridianDef;
                m->term1 -= m->preAdjCorr;
            } else
                m->preAdjCorr = 0.0;
        }
        // PartialD_Latitude_F / PartialD_Latitude (GEO:279-320): forward difference, increment 1e-4
        double lat = cart_to_lat(el, p1[0], p1[1], p1[2]);
        r.nst = 1;
        r.st[0] = s1;
        for (int k = 0; k < 3; ++k) {
            double c[3] = {p1[0], p1[1], p1[2]};
            c[k] += PRECISION_1E4;
            r.a[k] = (cart_to_lat(el, c[0], c[1], c[2]) - lat) / PRECISION_1E4;
        }
        *l_out = m->term1 - lat;
        return true;
    }
    case 'J':
    case 'Q': {
        if (first_run && m->measType == 'J') {
            if (std::fabs(st1.verticalDef) > E4_SEC_DEFLECTION) {
                m->preAdjCorr = st1.verticalDef / std::cos(st1.currentLatitude);
                m->term1 -= m->preAdjCorr;
            } else
                m->preAdjCorr = 0.0;
        }
        *l_out = m->term1 - st1.currentLongitude;
        double t = p1[0] * p1[1] / std::pow(p1[0] * p1[0] + p1[1] * p1[1], 1.5);
        r.nst = 1;
        r.st[0] = s1;
        r.a[0] = t * -1. / std::cos(st1.currentLongitude);
        r.a[1] = t / std::sin(st1.currentLongitude);
        r.a[2] = 0.;
        return true;
    }
    default:
        return false;
    }
}
