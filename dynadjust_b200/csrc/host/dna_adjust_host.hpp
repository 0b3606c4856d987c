// dna_adjust_host.hpp — host-side C++ mirror of the reference's `dna_adjust` interface for the solve path,
// implemented on the C-ABI (include/gadj.h).  Member names, argument meaning and error behaviour follow
// dynadjust/dynadjust/dnaadjust/dnaadjust.hpp:212-1362 (PrepareAdjustment :260, AdjustNetwork :405,
// GenerateStatistics :259, getters :336-354) and the wrapper's call order (dnaadjustwrapper.cpp:1142-1432);
// the text outputs follow dnaadjust_printer.cpp (header :3436-3599, iteration block :70-92, statistics :660-719,
// adjusted measurements, adjusted stations :3917-4070, positional uncertainty :2665-2770 / :4290-4470, station
// corrections :1349-1408 / :4146-4288).
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <cctype>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/gadj.h"
#include "../geodesy.h"
#include "dna_files.hpp"

namespace dynadjust_b200 {

enum ADJUST_MODE { SimultaneousMode = 0, PhasedMode = 1, Phased_Block_1Mode = 2 };
enum ADJUST_STATUS { ADJUST_SUCCESS = 0, ADJUST_MAX_ITERATIONS_EXCEEDED = 2, ADJUST_EXCEPTION_RAISED = 4 };

struct adjust_settings {            // the fields of project_settings.a / .g / .o the solve path reads
    std::string network_name;
    std::string input_folder = ".", output_folder = ".";
    int adjust_mode = SimultaneousMode;
    double iteration_threshold = (double)0.0005f;   // float in the reference (dnaoptions.hpp:432)
    uint32_t max_iterations = 10;
    double free_std_dev = 10.0, fixed_std_dev = 1.0e-6, confidence_interval = 95.0;
    bool scale_normals_to_unity = false;
    bool output_adj_msr = false;
    bool output_stn_blocks = false;        // --output-stn-blocks: phased modes print the station tables block by block
    bool output_msr_blocks = false;        // --output-msr-blocks: likewise the adjusted measurements
    bool output_pos_uncertainty = false;   // --output-pos-uncertainty: <net>.<mode>.apu
    bool output_corrections = false;       // --output-corrections-file: <net>.<mode>.cor
    bool export_sinex = false;             // --export-sinex-file: <net>[-block<k>].<frame>.snx with the dense block variance matrix
    bool apu_vcv_enu = false;              // --output-apu-vcv-units ENU (default XYZ)
    double hz_corr_threshold = 0.0, vt_corr_threshold = 0.0;   // dnaoptions.hpp:510
    bool update_binary_files = true;
    std::string type_b_global, type_b_file; // --type-b-sd-global "e,n,up" (metres, 1 sigma), --type-b-sd-file <file> (dnaoptions-interface.hpp)
    std::string station_constraints;       // --constraints "STN1,CCC,STN2,FFC" (dnaoptions.hpp:481)
    std::string command_line;
};

class dna_adjust {
  public:
    ~dna_adjust()
    {
        if (ctx_)
            gadj_destroy(ctx_);
    }

    // ---- PrepareAdjustment (ADJ:258): load .bst/.bms(/.seg), initialise, symbolic analysis, upload --------
    void PrepareAdjustment(const adjust_settings& s)
    {
        a_ = s;
        const std::string base = a_.input_folder + "/" + a_.network_name;
        bst_file_ = base + ".bst";
        bms_file_ = base + ".bms";
        dnafiles::load_binary(bst_file_, stn_, bst_meta_);
        dnafiles::load_binary(bms_file_, msr_, bms_meta_);
        ApplyConstraints();
        gadj_opts o;
        gadj_default_opts(&o);
        o.fixed_std_dev = a_.fixed_std_dev;
        o.free_std_dev = a_.free_std_dev;
        o.iteration_threshold = a_.iteration_threshold;
        o.max_iterations = a_.max_iterations;
        o.confidence_interval = a_.confidence_interval;
        o.scale_normals_to_unity = 1;   // internal equilibration is always safe; the flag is accepted for compatibility
        if (a_.export_sinex && a_.adjust_mode == SimultaneousMode) {
            // the SINEX file of a simultaneous adjustment carries the full variance matrix (PRN:2944-2946): one dense front
            if (stn_.size() > 12000)
                SignalExceptionAdjustment("--export-sinex-file in simultaneous mode needs the full variance matrix of the network "
                                          "(dense); segment the network (dnasegment) and run --phased-adjustment for per-block files");
            o.ordering = GADJ_ORDER_DENSE;
        }
        if (gadj_create(&o, &ctx_))
            SignalExceptionAdjustment(gadj_last_error(nullptr));
        check(gadj_set_stations(ctx_, stn_.data(), (uint32_t)stn_.size()));
        check(gadj_set_measurements(ctx_, msr_.data(), msr_.size()));
        if (a_.adjust_mode != SimultaneousMode) {
            dnafiles::load_seg(base + ".seg", seg_);
            std::vector<uint32_t> off{0}, isl;
            for (auto& b : seg_.isl) {
                isl.insert(isl.end(), b.begin(), b.end());
                off.push_back((uint32_t)isl.size());
            }
            check(gadj_set_blocks(ctx_, (uint32_t)seg_.isl.size(), off.data(), isl.data()));
        }
        check(gadj_prepare(ctx_));
        gadj_get_info(ctx_, &info_);
        apriori_llh_.resize(3 * stn_.size());
        apriori_xyz_.resize(3 * stn_.size());   // v_originalStations_ (ADJ:632-693)
        const gadj::Ellipsoid ell = gadj::make_ellipsoid(o.semi_major, o.inv_flattening);
        for (size_t i = 0; i < stn_.size(); ++i) {
            apriori_llh_[3 * i] = stn_[i].currentLatitude;
            apriori_llh_[3 * i + 1] = stn_[i].currentLongitude;
            apriori_llh_[3 * i + 2] = stn_[i].currentHeight;
            gadj::geo_to_cart(ell, stn_[i].currentLatitude, stn_[i].currentLongitude, stn_[i].currentHeight, &apriori_xyz_[3 * i]);
        }
    }

    // ---- AdjustNetwork (ADJ:2140) -> AdjustSimultaneous loop (ADJ:2413-2511) -------------------------------
    ADJUST_STATUS AdjustNetwork()
    {
        auto t0 = std::chrono::steady_clock::now();
        iterations_.clear();
        adjustStatus_ = ADJUST_SUCCESS;
        for (uint32_t i = 0; i < a_.max_iterations; ++i) {
            auto ti = std::chrono::steady_clock::now();
            gadj_iter_result r;
            check(gadj_iterate(ctx_, i == 0 ? GADJ_ITER_NORMALS : 0, &r));
            r.ms_inverse = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - ti).count();  // wall
            iterations_.push_back(r);
            maxCorr_ = r.max_corr;
            if (std::fabs(r.max_corr) <= a_.iteration_threshold)
                break;
        }
        if (iterations_.size() == a_.max_iterations && std::fabs(maxCorr_) > a_.iteration_threshold)
            adjustStatus_ = ADJUST_MAX_ITERATIONS_EXCEEDED;   // ADJ:2523-2525
        check(gadj_form_inverse(ctx_));                        // rigorous variances (v_rigorousVariances_)
        total_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return adjustStatus_;
    }

    // ---- GenerateStatistics (ADJ:6802) ------------------------------------------------------------------------
    void GenerateStatistics()
    {
        check(gadj_statistics(ctx_, &stats_, 1));   // also refreshes the records: adjusted lat/lon/h and measurement statistics
        est_.resize(3 * stn_.size());
        vcv_.resize(9 * stn_.size());
        check(gadj_get_estimates(ctx_, est_.data()));
        check(gadj_get_station_vcvs(ctx_, vcv_.data()));
        ApplyTypeBUncertainties();
        ComputeTestStat();
    }

    // getters (ADJH:336-354)
    double GetChiSquared() const { return stats_.chi_squared; }
    double GetSigmaZero() const { return stats_.sigma_zero; }
    int64_t GetDegreesOfFreedom() const { return stats_.dof; }
    uint32_t GetMeasurementCount() const { return stats_.measurement_params; }
    uint32_t GetUnknownsCount() const { return stats_.unknown_params; }
    double GetGlobalPelzerRel() const { return stats_.global_pelzer; }
    double GetChiSquaredUpperLimit() const { return chiUpper_; }
    double GetChiSquaredLowerLimit() const { return chiLower_; }
    uint32_t CurrentIteration() const { return (uint32_t)iterations_.size(); }
    ADJUST_STATUS GetStatus() const { return adjustStatus_; }
    const gadj_info& Info() const { return info_; }

    std::string ModeSuffix() const
    {   // output naming (WRAP:659-734)
        switch (a_.adjust_mode) {
        case PhasedMode: return "phased";
        case Phased_Block_1Mode: return "phased-block1";
        default: return "simult";
        }
    }

    // ---- outputs (WRAP:1397-1432) --------------------------------------------------------------------------------
    void PrintAdjustedNetwork()
    {
        const std::string stem = a_.output_folder + "/" + a_.network_name + "." + ModeSuffix();
        std::ofstream adj(stem + ".adj");
        PrintOutputFileHeaderInfo(adj, "DYNADJUST ADJUSTMENT OUTPUT FILE", stem + ".adj");
        adj << "\n+ Initialising adjustment\n+ Loading network files\n+ Allocating memory\n\n+ Preparing for adjustment...  done.\n";
        adj << "+ Commencing " << (a_.adjust_mode == SimultaneousMode ? "simultaneous" : "phased") << " adjustment\n\n";
        for (size_t i = 0; i < iterations_.size(); ++i)
            PrintIteration(adj, (uint32_t)i + 1, iterations_[i]);
        PrintStatistics(adj);
        if (a_.output_adj_msr)
            PrintAdjustedNetworkMeasurements(adj);
        std::ofstream xyz(stem + ".xyz");
        PrintOutputFileHeaderInfo(xyz, "DYNADJUST COORDINATE OUTPUT FILE", stem + ".xyz");
        PrintAdjustedNetworkStations(adj, xyz);
        if (a_.output_pos_uncertainty)
            PrintPositionalUncertainty(stem + ".apu");
        if (a_.output_corrections)
            PrintNetworkStationCorrections(stem + ".cor");
        if (a_.export_sinex)
            PrintEstimatedStationCoordinatestoSNX();
    }

    // ---- .snx (PrintEstimatedStationCoordinatestoSNX PRN:2906-3010, DnaIoSnx::SerialiseSinex snx_file_writer.cpp) -----------
    // One file per block, <net>-block<k>.<frame>.snx (phased; block-1 mode: the first only), or <net>.<frame>.snx
    // (simultaneous): SITE/ID, SOLUTION/STATISTICS, SOLUTION/ESTIMATE and the lower triangle of the block's dense
    // variance matrix, SOLUTION/MATRIX_ESTIMATE L COVA.
    void PrintEstimatedStationCoordinatestoSNX()
    {
        const uint32_t nblocks = (uint32_t)info_.nfronts;
        const bool phased = a_.adjust_mode != SimultaneousMode;
        const std::string frame = frame_name();
        for (uint32_t b = 0; b < nblocks; ++b) {
            if (a_.adjust_mode == Phased_Block_1Mode && b > 0)
                break;
            uint32_t n = 0;
            check(gadj_get_block_vcv(ctx_, b, &n, nullptr, 0, nullptr));
            std::vector<uint32_t> st(n);
            const size_t dim = 3 * (size_t)n;
            std::vector<double> q(dim * (dim + 1) / 2);
            check(gadj_get_block_vcv(ctx_, b, &n, st.data(), n, q.data()));
            std::string file = a_.output_folder + "/" + a_.network_name;
            if (phased)
                file += "-block" + std::to_string(b + 1);
            file += "." + frame + ".snx";
            std::ofstream os(file);
            auto at = [&](size_t i, size_t j) { return i >= j ? q[j * dim - j * (j - 1) / 2 + (i - j)] : q[i * dim - i * (i - 1) / 2 + (j - i)]; };
            const std::string line = "*-------------------------------------------------------------------------------";
            char buf[256];
            const std::string epoch = sinex_date(bst_meta_.epoch, false), now = sinex_date("", true);
            snprintf(buf, sizeof(buf), "%%=SNX 2.00 DNA %s DNA %s %s P %05u 0 S           ", now.c_str(), epoch.c_str(), epoch.c_str(),
                     (unsigned)stats_.unknown_params);
            os << buf << "\n" << line << "\n+FILE/REFERENCE\n"
               << "*INFO_TYPE_________ INFO________________________________________________________\n"
               << " DESCRIPTION        Network " << a_.network_name << "\n";
            std::ostringstream what;
            if (nblocks > 1)
                what << "Phased adjustment results. Block " << b + 1 << " of " << nblocks;
            else
                what << "Simultaneous adjustment results.";
            os << " OUTPUT             " << std::left << std::setw(60) << what.str() << "\n"
               << " SOFTWARE           b200-geodetic-adjust 0.1 (libgadj, sm_100a)\n"
               << " INPUT              " << std::left << std::setw(60) << bst_file_ << "\n"
               << " INPUT              " << std::left << std::setw(60) << bms_file_ << "\n-FILE/REFERENCE\n" << line << "\n+FILE/COMMENT\n";
            if (nblocks > 1)
                os << " This file contains the rigorous estimates for block " << b + 1 << " of a segmented\n network comprised of " << nblocks
                   << " blocks. Due to the way in which junction stations\n are carried through successive blocks, stations appearing in this "
                      "file\n may also be found in other SINEX files relating to this network, such as\n "
                   << a_.network_name << "-block1.snx, " << a_.network_name << "-block2.snx, etc.\n";
            os << "-FILE/COMMENT\n" << line << "\n+SITE/ID\n"
               << "*CODE PT __DOMES__ T _STATION DESCRIPTION__ APPROX_LON_ APPROX_LAT_ _APP_H_\n";
            for (uint32_t i = 0; i < n; ++i) {
                const dna_stn_t& s = stn_[st[i]];
                const std::string name = s.stationName, desc = s.description;
                snprintf(buf, sizeof(buf), " %-4s %2s %-9s %1s %-22s %11s %11s %7.1f", name.substr(0, 4).c_str(), "A", name.substr(0, 9).c_str(), "P",
                         desc.substr(0, 22).c_str(), dms_spaced5(s.currentLongitude).c_str(), dms_spaced5(s.currentLatitude).c_str(),
                         s.currentHeight);
                os << buf << "\n";
            }
            os << "-SITE/ID\n" << line << "\n+SOLUTION/STATISTICS\n*_STATISTICAL PARAMETER________ __VALUE(S)____________\n";
            snprintf(buf, sizeof(buf), " %-30s %22u\n %-30s %22u\n %-30s %22lld\n %-30s %22.6f\n", "NUMBER OF OBSERVATIONS",
                     (unsigned)stats_.measurement_params, "NUMBER OF UNKNOWNS", (unsigned)stats_.unknown_params, "NUMBER OF DEGREES OF FREEDOM",
                     (long long)stats_.measurement_params - (long long)stats_.unknown_params, "VARIANCE FACTOR", stats_.sigma_zero);
            os << buf << "-SOLUTION/STATISTICS\n" << line << "\n+SOLUTION/ESTIMATE\n"
               << "*INDEX TYPE__ CODE PT SOLN _REF_EPOCH__ UNIT S __ESTIMATED VALUE____ _STD_DEV___\n";
            unsigned index = 1;
            for (uint32_t i = 0; i < n; ++i)
                for (int c = 0; c < 3; ++c) {
                    const std::string name = stn_[st[i]].stationName;
                    char val[40], sd[40];
                    snprintf(val, sizeof(val), "%.14E", est_[3 * (size_t)st[i] + c]);
                    snprintf(sd, sizeof(sd), "%.5E", std::sqrt(std::fabs(at(3 * i + c, 3 * i + c))));
                    snprintf(buf, sizeof(buf), " %5u STA%c   %-4s %2s 0001 %s %-4s 0 %21s %11s", index++, "XYZ"[c], name.substr(0, 4).c_str(), "A",
                             epoch.c_str(), "m", val, sd);
                    os << buf << "\n";
                }
            os << "-SOLUTION/ESTIMATE\n" << line << "\n+SOLUTION/MATRIX_ESTIMATE L COVA\n"
               << "*PARA1 PARA2 ____PARA2+0__________ ____PARA2+1__________ ____PARA2+2__________\n";
            for (size_t row = 0; row < dim; ++row) {
                int field = 1;
                bool fresh = true;
                for (size_t col = 0; col <= row; ++col) {
                    if (fresh) {
                        snprintf(buf, sizeof(buf), " %5zu %5zu ", row + 1, col + 1);
                        os << buf;
                        fresh = false;
                    }
                    snprintf(buf, sizeof(buf), "%21.14E ", at(row, col));
                    os << buf;
                    if (row == col || ++field > 3) {
                        os << "\n";
                        fresh = true;
                        field = 1;
                    }
                }
            }
            os << "-SOLUTION/MATRIX_ESTIMATE L COVA\n%ENDSNX\n";
        }
    }

    // PrintAdjustedNetworkStations (PRN:535-595): one list of every station; in the phased modes with
    // --output-stn-blocks one table per block (inner + junction stations); block-1 mode stops after the first block.
    void PrintAdjustedNetworkStations(std::ostream& adj, std::ostream& xyz) const
    {
        const bool phased = a_.adjust_mode != SimultaneousMode && !seg_.isl.empty();
        if (!phased || (!a_.output_stn_blocks && a_.adjust_mode != Phased_Block_1Mode)) {
            PrintAdjStations(adj, nullptr);
            PrintAdjStations(xyz, nullptr);
            return;
        }
        for (size_t b = 0; b < seg_.isl.size(); ++b) {
            std::vector<uint32_t> list(seg_.isl[b]);
            if (b < seg_.jsl.size())
                list.insert(list.end(), seg_.jsl[b].begin(), seg_.jsl[b].end());
            std::sort(list.begin(), list.end());
            list.erase(std::unique(list.begin(), list.end()), list.end());
            if (a_.output_stn_blocks) {
                adj << "\nBlock " << b + 1 << "\n";
                xyz << "\nBlock " << b + 1 << "\n";
            }
            PrintAdjStations(adj, &list);
            PrintAdjStations(xyz, &list);
            if (a_.adjust_mode == Phased_Block_1Mode)
                break;   // only the first block is reported (PRN:586-588)
        }
    }

    // ---- .apu (PrintPositionalUncertainty PRN:2665-2770, PrintPosUncertainty PRN:4326-4432) -------------------------
    // Per station: horizontal / vertical positional uncertainty at 95 %, 1-sigma error ellipse, and the upper triangle
    // of its 3x3 variance block (XYZ or ENU).  Stations are listed as one block (the reference's layout for
    // simultaneous adjustments and for phased ones without --output-stn-blocks).
    void PrintPositionalUncertainty(const std::string& file) const
    {
        std::ofstream os(file);
        PrintStationFileHeader(os, "POSITIONAL UNCERTAINTY", file);
        auto var = [&](const char* n, const std::string& v) { os << std::left << std::setw(35) << n << v << "\n"; };
        var("PU confidence interval:", "95.0%");
        var("Error ellipse axes:", "68.3% (1 sigma)");
        var("Variances:", "68.3% (1 sigma)");
        var("Stations printed in blocks:", "No");
        var("Variance matrix units:", a_.apu_vcv_enu ? "ENU" : "XYZ");
        var("Full covariance matrix:", "No");
        if (!a_.type_b_global.empty())
            var("Type B uncertainties:", a_.type_b_global);
        if (!a_.type_b_file.empty())
            var("Type B uncertainty file:", a_.type_b_file);
        os << std::string(80, '-') << "\n\n";
        os << "Positional uncertainty of adjusted station coordinates\n";
        os << "------------------------------------------------------\n\n";
        char buf[512];
        const char* vn = a_.apu_vcv_enu ? "enu" : "XYZ";
        char v1[16], v2[16], v3[16];
        snprintf(v1, sizeof(v1), "Variance(%c)", vn[0]);
        snprintf(v2, sizeof(v2), "Variance(%c)", vn[1]);
        if (a_.apu_vcv_enu)
            snprintf(v3, sizeof(v3), "Variance(up)");
        else
            snprintf(v3, sizeof(v3), "Variance(Z)");
        snprintf(buf, sizeof(buf), "%-20s%2s%14s%15s%11s%11s%13s%13s%13s%19s%19s%19s", "Station", "", "Latitude", "Longitude", "Hz PosU",
                 "Vt PosU", "Semi-major", "Semi-minor", "Orientation", v1, v2, v3);
        os << buf << "\n" << std::string(20 + 2 + 14 + 15 + 11 + 11 + 13 + 13 + 13 + 19 + 19 + 19, '-') << "\n";
        const int pad = 20 + 2 + 14 + 15 + 11 + 11 + 13 + 13 + 13;
        for (size_t i = 0; i < stn_.size(); ++i) {
            const dna_stn_t& s = stn_[i];
            const double* q = &vcv_[9 * i];
            double ql[9];
            to_local(q, s.currentLatitude, s.currentLongitude, ql);
            double smaj, smin, az, hz, vt;
            ErrorEllipseParameters(ql, smaj, smin, az);
            PositionalUncertainty(smaj, smin, std::sqrt(std::fabs(ql[8])), hz, vt);
            const double* v = a_.apu_vcv_enu ? ql : q;
            snprintf(buf, sizeof(buf), "%-20s%2s%14.9f%15.9f%11.4f%11.4f%13.4f%13.4f%13.4f%19.9e%19.9e%19.9e", s.stationName, "",
                     rad_to_dms(s.currentLatitude), rad_to_dms(s.currentLongitude), hz, vt, smaj, smin, rad_to_dms(az), v[0], v[1], v[2]);
            os << buf << "\n";
            snprintf(buf, sizeof(buf), "%*s%19.9e%19.9e", pad + 19, "", v[4], v[5]);
            os << buf << "\n";
            snprintf(buf, sizeof(buf), "%*s%19.9e", pad + 38, "", v[8]);
            os << buf << "\n";
        }
    }

    // ---- .cor (PrintNetworkStationCorrections PRN:1349-1408, PrintCorStation PRN:4146-4230) -----------------------------
    // Per station: azimuth, vertical angle, slope and horizontal distance of the shift a-priori -> adjusted position and
    // its local e / n / up components; stations inside both thresholds are left out.
    void PrintNetworkStationCorrections(const std::string& file) const
    {
        std::ofstream os(file);
        PrintStationFileHeader(os, "CORRECTIONS", file);
        os << std::left << std::setw(35) << "Stations printed in blocks:" << "No\n" << std::string(80, '-') << "\n\n";
        os << "Corrections to stations\n------------------------------------------\n\n";
        char buf[512];
        snprintf(buf, sizeof(buf), "%-20s%2s%19s%19s%19s%19s%11s%11s%11s", "Station", "", "Azimuth", "V. Angle", "S. Distance", "H. Distance",
                 "east", "north", "up");
        os << buf << "\n" << std::string(20 + 2 + 4 * 19 + 3 * 11, '-') << "\n";
        for (size_t i = 0; i < stn_.size(); ++i) {
            const dna_stn_t& s = stn_[i];
            const double d[3] = {est_[3 * i] - apriori_xyz_[3 * i], est_[3 * i + 1] - apriori_xyz_[3 * i + 1],
                                 est_[3 * i + 2] - apriori_xyz_[3 * i + 2]};
            const double lat = s.currentLatitude, lon = s.currentLongitude;   // the adjusted position, as in the reference
            const double e = -std::sin(lon) * d[0] + std::cos(lon) * d[1];
            const double n = -std::sin(lat) * std::cos(lon) * d[0] - std::sin(lat) * std::sin(lon) * d[1] + std::cos(lat) * d[2];
            const double u = std::cos(lat) * std::cos(lon) * d[0] + std::cos(lat) * std::sin(lon) * d[1] + std::sin(lat) * d[2];
            const bool tiny = std::fabs(e) < 1e-5 && std::fabs(n) < 1e-5;
            double va = std::atan2(u, std::sqrt(e * e + n * n));
            if (tiny && std::fabs(u) < 1e-5)
                va = 0.0;
            if (std::fabs(u) < a_.vt_corr_threshold)
                continue;
            const double hd = std::sqrt(e * e + n * n);
            if (hd < a_.hz_corr_threshold)
                continue;
            double az = tiny ? 0.0 : direction_en(e, n);
            const double sd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            snprintf(buf, sizeof(buf), "%-20s%2s%19s%19s%19.4f%19.4f%11.4f%11.4f%11.4f", s.stationName, "",
                     FormatDmsString(rad_to_dms(az)).c_str(), FormatDmsString(rad_to_dms(va)).c_str(), sd, hd, e, n, u);
            os << buf << "\n";
        }
        os << "\n";
    }

    // UpdateBinaryFiles (ADJ:445-470): adjusted coordinates / statistics back to .bst/.bms with reduced = true
    void UpdateBinaryFiles()
    {
        bst_meta_.reduced = true;
        bms_meta_.reduced = true;
        snprintf(bst_meta_.modifiedBy, sizeof(bst_meta_.modifiedBy), "%s", "adjust");
        snprintf(bms_meta_.modifiedBy, sizeof(bms_meta_.modifiedBy), "%s", "adjust");
        dnafiles::write_binary(bst_file_, stn_, bst_meta_);
        dnafiles::write_binary(bms_file_, msr_, bms_meta_);
    }

  private:
    // Type B uncertainties (LoadTypeBUncertainties ADJ:10231-10323, dnaiotbu.cpp, PrintAdjStation PRN:4000-4029): 1-sigma
    // east / north / up values in metres — one set for every station on the command line, site-specific ones from a
    // "!#=DNA 1.00 TBU" file (station name in the first 20 columns) taking precedence — are added, as variances rotated
    // into the Cartesian frame at the station's estimated position, to the station variance blocks that every report
    // (.adj, .xyz, .apu) reads.
    void ApplyTypeBUncertainties()
    {
        if (a_.type_b_global.empty() && a_.type_b_file.empty())
            return;
        auto parse3 = [&](const std::vector<std::string>& t, const std::string& what, double* enu) {
            std::vector<double> v;
            for (const std::string& x : t) {
                char* end = nullptr;
                const double d = std::strtod(x.c_str(), &end);
                if (x.empty() || end == x.c_str() || *end != 0)
                    SignalExceptionAdjustment("  Type b uncertainty '" + x + "' is not a number:\n    " + what);
                v.push_back(d);
            }
            enu[0] = enu[1] = enu[2] = 0.0;
            if (v.size() >= 3) {
                enu[0] = v[0] * v[0];
                enu[1] = v[1] * v[1];
                enu[2] = v[2] * v[2];
            } else if (v.size() == 2) {       // east, north
                enu[0] = v[0] * v[0];
                enu[1] = v[1] * v[1];
            } else if (v.size() == 1)         // up
                enu[2] = v[0] * v[0];
            else
                SignalExceptionAdjustment("  No Type b uncertainties provided:\n    " + what);
        };
        std::vector<double> tb(3 * stn_.size(), 0.0);
        std::vector<char> has(stn_.size(), 0);
        if (!a_.type_b_global.empty()) {
            std::vector<std::string> tok;
            std::stringstream ss(a_.type_b_global);
            for (std::string t; std::getline(ss, t, ',');)
                tok.push_back(t);
            double enu[3];
            parse3(tok, a_.type_b_global, enu);
            for (size_t i = 0; i < stn_.size(); ++i) {
                std::copy(enu, enu + 3, &tb[3 * i]);
                has[i] = 1;
            }
        }
        if (!a_.type_b_file.empty()) {
            std::ifstream f(a_.type_b_file);
            std::string line;
            if (!f || !std::getline(f, line))
                SignalExceptionAdjustment("load_tbu_file(): An error was encountered when opening " + a_.type_b_file + ".");
            if (line.size() < 15 || line.compare(0, 6, "!#=DNA") != 0 || (line.substr(12, 3) != "TBU" && line.substr(12, 3) != "tbu"))
                SignalExceptionAdjustment("  The supplied filetype is not recognised:\n  " + line);
            std::unordered_map<std::string, uint32_t> by_name;
            for (size_t i = 0; i < stn_.size(); ++i)
                by_name.emplace(stn_[i].stationName, (uint32_t)i);
            while (std::getline(f, line)) {
                if (line.empty() || line[0] == '*' || line.find_first_not_of(" \t\r") == std::string::npos)
                    continue;
                std::string name = line.substr(0, 20);
                name.erase(name.find_last_not_of(" \t\r") + 1);
                auto it = by_name.find(name);
                if (it == by_name.end() || line.size() <= 20)
                    continue;                 // stations outside the network are ignored (dnaiotbu.cpp:255-262)
                std::vector<std::string> tok;
                std::stringstream ss(line.substr(20));
                for (std::string t; ss >> t;)
                    tok.push_back(t);
                if (tok.empty())
                    continue;
                tok.resize(3, "0");
                double enu[3];
                parse3(tok, line, enu);
                std::copy(enu, enu + 3, &tb[3 * (size_t)it->second]);
                has[it->second] = 1;
            }
        }
        for (size_t i = 0; i < stn_.size(); ++i) {
            if (!has[i])
                continue;
            const double lat = stn_[i].currentLatitude, lon = stn_[i].currentLongitude;
            const double sl = std::sin(lat), cl = std::cos(lat), so = std::sin(lon), co = std::cos(lon);
            const double R[3][3] = {{-so, -sl * co, cl * co}, {co, -sl * so, cl * so}, {0, cl, sl}};   // local -> Cartesian
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    vcv_[9 * i + 3 * a + b] += R[a][0] * tb[3 * i] * R[b][0] + R[a][1] * tb[3 * i + 1] * R[b][1] + R[a][2] * tb[3 * i + 2] * R[b][2];
        }
    }

    // NetworkDataLoader::ApplyConstraints (network_data_loader.cpp:211-263): user-supplied "station,constraint" pairs
    // override the constraints in the .bst records.  The reference looks the names up in <net>.map (names sorted, binary
    // search); the same names are in the station records, so they are indexed here directly.
    void ApplyConstraints()
    {
        if (a_.station_constraints.empty())
            return;
        std::vector<std::string> tok;
        std::stringstream ss(a_.station_constraints);
        for (std::string t; std::getline(ss, t, ',');) {
            size_t b = t.find_first_not_of(" \t"), e = t.find_last_not_of(" \t");
            tok.push_back(b == std::string::npos ? std::string() : t.substr(b, e - b + 1));
        }
        std::unordered_map<std::string, uint32_t> by_name;
        for (size_t i = 0; i < stn_.size(); ++i)
            by_name.emplace(stn_[i].stationName, (uint32_t)i);
        for (size_t k = 0; k + 1 < tok.size(); k += 2) {
            std::string c = tok[k + 1];
            for (char& ch : c)
                ch = (char)std::toupper((unsigned char)ch);
            auto it = by_name.find(tok[k]);
            if (it == by_name.end())
                SignalExceptionAdjustment("The supplied constraint station '" + tok[k] + "' is not in the stations map");
            if (c.size() != 3 || c.find_first_not_of("CF") != std::string::npos)   // CDnaStation::IsValidConstraint
                SignalExceptionAdjustment("Invalid station constraint: '" + tok[k + 1] + "'");
            snprintf(stn_[it->second].stationConst, sizeof(stn_[it->second].stationConst), "%s", c.c_str());
        }
    }

    void check(int rc)
    {
        if (rc)
            SignalExceptionAdjustment(gadj_last_error(ctx_));
    }
    [[noreturn]] void SignalExceptionAdjustment(const std::string& msg)
    {   // ADJ:10049-10069
        adjustStatus_ = ADJUST_EXCEPTION_RAISED;
        throw std::runtime_error(msg);
    }

    // regularised lower incomplete gamma P(a, x) (series / continued fraction) for the chi-square limits that the
    // reference takes from boost::math (ADJ:6866-6911)
    static double gamma_p(double a, double x)
    {
        if (x <= 0)
            return 0;
        const double gln = std::lgamma(a);
        if (x < a + 1) {
            double ap = a, sum = 1 / a, del = sum;
            for (int n = 0; n < 100000; ++n) {
                ap += 1;
                del *= x / ap;
                sum += del;
                if (std::fabs(del) < std::fabs(sum) * 1e-16)
                    break;
            }
            return sum * std::exp(-x + a * std::log(x) - gln);
        }
        double b = x + 1 - a, c = 1 / 1e-300, d = 1 / b, h = d;
        for (int i = 1; i < 100000; ++i) {
            double an = -i * (i - a);
            b += 2;
            d = an * d + b;
            if (std::fabs(d) < 1e-300)
                d = 1e-300;
            c = b + an / c;
            if (std::fabs(c) < 1e-300)
                c = 1e-300;
            d = 1 / d;
            double del = d * c;
            h *= del;
            if (std::fabs(del - 1) < 1e-16)
                break;
        }
        return 1 - std::exp(-x + a * std::log(x) - gln) * h;
    }
    static double chi2_quantile(double p, double dof)
    {
        // Wilson-Hilferty start, bisection/Newton polish on the CDF
        double z = inv_norm(p);
        double t = 1 - 2 / (9 * dof) + z * std::sqrt(2 / (9 * dof));
        double x = dof * t * t * t;
        double lo = 0, hi = std::max(4 * dof, x * 4 + 100);
        for (int it = 0; it < 200; ++it) {
            double f = gamma_p(dof / 2, x / 2) - p;
            if (f > 0)
                hi = x;
            else
                lo = x;
            double pdf = std::exp((dof / 2 - 1) * std::log(x / 2) - x / 2 - std::lgamma(dof / 2)) / 2;
            double xn = pdf > 0 ? x - f / pdf : 0.5 * (lo + hi);
            if (!(xn > lo && xn < hi))
                xn = 0.5 * (lo + hi);
            if (std::fabs(xn - x) < 1e-12 * x)
                return xn;
            x = xn;
        }
        return x;
    }
    static double inv_norm(double p)
    {
        // Acklam
        static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                   1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
        static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                   6.680131188771972e+01, -1.328068155288572e+01};
        static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                   -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
        static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
        double q, r;
        if (p < 0.02425) {
            q = std::sqrt(-2 * std::log(p));
            return (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
                   ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
        }
        if (p <= 1 - 0.02425) {
            q = p - 0.5;
            r = q * q;
            return (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
                   (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
        }
        q = std::sqrt(-2 * std::log(1 - p));
        return -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
               ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1);
    }

    // ComputeTestStat (ADJ:6866-6911)
    void ComputeTestStat()
    {
        double conf = (100. - a_.confidence_interval) * 0.01 * 0.5;
        double dof = (double)stats_.dof;
        if (dof <= 0) {
            chiUpper_ = chiLower_ = 0;
            passFail_ = 2;
            return;
        }
        chiUpper_ = chi2_quantile(1 - conf, dof) / dof;
        chiLower_ = chi2_quantile(conf, dof) / dof;
        passFail_ = stats_.sigma_zero < chiLower_ ? 1 : (stats_.sigma_zero > chiUpper_ ? 2 : 0);
    }

    // RadtoDms / DegtoDms (dnatemplatecalcfuncs.hpp:206-222, 283-288): ddd.mmssss as a number
    static double rad_to_dms(double rad)
    {
        const double deg = rad * 180.0 / 3.14159265358979323846;
        double v = std::fabs(deg);
        const double d = std::floor(v);
        double m = std::floor((v - d) * 60.0);
        double s = (v - d - m / 60.0) * 3600.0;
        if (std::fabs(s - 60.0) < 0.000000001) {
            s = 0.0;
            m += 1.0;
        }
        v = d + m / 100.0 + s / 10000.0;
        return deg < 0.0 ? -v : v;
    }
    // FormatDmsString(dms, 4, withSpaces, no symbols) (dnatemplatefuncs.hpp:253-310): "ddd mm ss"
    static std::string FormatDmsString(double dms)
    {
        char b[64];
        snprintf(b, sizeof(b), "%.4f", dms);
        std::string t(b);
        size_t dot = t.find('.');
        if (dot == std::string::npos)
            return t;
        t.replace(dot, 1, " ");
        t.insert(dot + 3, " ");
        return t;
    }
    // atan_2 / Direction (dnatemplatecalcfuncs.hpp:350-362, dnatemplategeodesyfuncs.hpp:679-693)
    static double atan_2(double x, double y)
    {
        const double t = std::atan(x / y);
        if (y < 0)
            return t + 3.14159265358979323846;
        return x > 0 ? t : t + 2 * 3.14159265358979323846;
    }
    static double direction_en(double e, double n)
    {
        double d = std::fabs(e) < std::fabs(n) ? atan_2(e, n) : 3.14159265358979323846 / 2 - atan_2(n, e);
        if (d < 0)
            d += 2 * 3.14159265358979323846;
        return d;
    }
    // V_local = R^T V_cart R with R = local (e, n, up) -> Cartesian at (lat, lon)  (PropagateVariances_LocalCart, MFN:592-621)
    static void to_local(const double* q, double lat, double lon, double* out)
    {
        const double sl = std::sin(lat), cl = std::cos(lat), so = std::sin(lon), co = std::cos(lon);
        const double R[3][3] = {{-so, -sl * co, cl * co}, {co, -sl * so, cl * so}, {0, cl, sl}};
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double v = 0;
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j)
                        v += R[i][a] * q[3 * i + j] * R[j][b];
                out[3 * a + b] = v;
            }
    }
    // ErrorEllipseParameters (MFN:840-891) on the local e / n block
    static void ErrorEllipseParameters(const double* ql, double& smaj, double& smin, double& az)
    {
        smaj = smin = az = -1.;
        const double e2 = ql[0], n2 = ql[4], en = ql[1];
        double W = (e2 - n2) * (e2 - n2) + 4. * en * en;
        if (W < 0.0) {
            if (std::fabs(W) > 1e-15)
                return;
            W = 0.0;
        }
        const double a2 = 0.5 * (e2 + n2 + std::sqrt(W)), b2 = 0.5 * (e2 + n2 - std::sqrt(W));
        if (a2 < 0.0 || b2 < 0.0)
            return;
        smaj = std::sqrt(a2);
        smin = std::sqrt(b2);
        if (std::fabs(e2 - n2) < 1e-25)
            az = en < 1e-25 ? 0. : 3.14159265358979323846 / 4.;
        else
            az = 0.5 * atan_2(en + en, n2 - e2);
    }
    // PositionalUncertainty (MFN:808-826; coefficients dnaconsts.hpp:105-108)
    static void PositionalUncertainty(double smaj, double smin, double sd_ht, double& hz, double& vt)
    {
        hz = vt = -1.;
        if (smaj < 0.0 || smin < 0.0)
            return;
        const double c = smin / smaj;
        hz = smaj * (1.96079 + 0.004071 * c + 0.114276 * c * c + 0.371625 * c * c * c);
        vt = sd_ht * 1.96;
    }
    // print_file_header + file name (dnaiostreamfuncs.hpp:115-141, PRN:1156-1162)
    void PrintStationFileHeader(std::ostream& os, const char* type, const std::string& file) const
    {
        auto var = [&](const char* n, const std::string& v) { os << std::left << std::setw(35) << n << v << "\n"; };
        os << std::string(80, '-') << "\nDYNADJUST " << type << " OUTPUT FILE\n\n";
        var("Version:", "b200-geodetic-adjust 0.1 (libgadj, sm_100a)");
        var("Build:", std::string(__DATE__) + ", " + __TIME__);
        var("File name:", file);
        os << "\n";
    }

    static std::string hp_dms(double rad)
    {   // degrees.minutes-seconds "HP" notation, 9 decimals (FormatDmsString / RadtoDms)
        double deg = rad * 180.0 / 3.14159265358979323846;
        double sgn = deg < 0 ? -1 : 1;
        deg = std::fabs(deg);
        double d = std::floor(deg + 1e-13);
        double m = std::floor((deg - d) * 60 + 1e-11);
        double s = ((deg - d) * 60 - m) * 60;
        if (s < 0)
            s = 0;
        if (s >= 59.999995) {
            s = 0;
            m += 1;
        }
        if (m >= 60) {
            m -= 60;
            d += 1;
        }
        char buf[64];
        snprintf(buf, sizeof(buf), "%.9f", sgn * (d + m / 100.0 + s / 10000.0));
        return buf;
    }

    void PrintOutputFileHeaderInfo(std::ostream& os, const char* title, const std::string& file) const
    {   // PRN:3436-3599
        const std::string dash(80, '-');
        auto var = [&](const char* n, const std::string& v) { os << std::left << std::setw(35) << n << v << "\n"; };
        os << dash << "\n" << title << "\n\n";
        var("Version:", "b200-geodetic-adjust 0.1 (libgadj, sm_100a)");
        var("Build:", std::string(__DATE__) + ", " + __TIME__);
        var("File name:", file);
        os << "\n";
        var("Command line arguments:", a_.command_line);
        os << "\n";
        var("Stations file:", bst_file_);
        var("Measurements file:", bms_file_);
        var("Reference frame:", std::string("EPSG ") + bst_meta_.epsgCode);
        var("Epoch:", bst_meta_.epoch);
        std::ostringstream t;
        t << a_.fixed_std_dev;
        if (!a_.station_constraints.empty())
            var("Station constraints:", a_.station_constraints);   // PRN:3490
        var("Constrained Station S.D. (m):", t.str());
        t.str("");
        t << a_.free_std_dev;
        var("Free Station S.D. (m):", t.str());
        t.str("");
        t << (float)a_.iteration_threshold;
        var("Iteration threshold:", t.str());
        var("Maximum iterations:", std::to_string(a_.max_iterations));
        t.str("");
        t << std::fixed << std::setprecision(1) << a_.confidence_interval << "%";
        var("Test confidence interval:", t.str());
        var("Uncertainties SD(e,n,up):", "68.3% (1 sigma)");
        var("Station coordinate types:", "PLHhXYZ");
        var("Stations printed in blocks:", "No");
        t.str("");
        t << info_.nfronts << " fronts on " << info_.nlevels << " levels (supernodal Cholesky on B200)";
        var("Elimination tree:", t.str());
        os << dash << "\n";
    }

    void PrintIteration(std::ostream& os, uint32_t it, const gadj_iter_result& r) const
    {   // PRN:70-92, OutputLargestCorrection ADJ:7357-7448
        const std::string dash(80, '-');
        os << "\n" << dash << "\n" << std::left << std::setw(35) << "ITERATION" << it << "\n\n";
        char buf[64];
        double sec = r.ms_inverse / 1e3;
        snprintf(buf, sizeof(buf), "00:00:%09.6f", sec);
        os << std::left << std::setw(35) << "Elapsed time" << buf << "\n";
        const dna_stn_t& s = stn_[r.max_corr_station];
        os << std::left << std::setw(35) << "Maximum station correction" << "Station " << s.stationName << "\n";
        // Rotate_CartLocal at the station's a-priori geographic position
        double lat = apriori_llh_[3 * r.max_corr_station], lon = apriori_llh_[3 * r.max_corr_station + 1];
        const double* d = r.max_corr_xyz;
        double e = -std::sin(lon) * d[0] + std::cos(lon) * d[1];
        double n = -std::sin(lat) * std::cos(lon) * d[0] - std::sin(lat) * std::sin(lon) * d[1] + std::cos(lat) * d[2];
        double u = std::cos(lat) * std::cos(lon) * d[0] + std::cos(lat) * std::sin(lon) * d[1] + std::sin(lat) * d[2];
        double big = std::max(std::fabs(e), std::max(std::fabs(n), std::fabs(u)));
        os << std::setw(35) << " ";
        if (big > 0.000999)
            os << std::fixed << std::setprecision(3) << e << ", " << n << ", " << u;
        else if (big > 0.00009)
            os << std::fixed << std::setprecision(4) << e << ", " << n << ", " << u;
        else
            os << std::scientific << std::setprecision(1) << e << ", " << n << ", " << u;
        os << " (e, n, up)\n\n";
        os.unsetf(std::ios::floatfield);
    }

    void PrintStatistics(std::ostream& os) const
    {   // PRN:660-719
        const std::string dash(80, '-');
        auto var = [&](const char* n) -> std::ostream& { return os << std::left << std::setw(35) << n; };
        os << "\n" << dash << "\n";
        var("SOLUTION") << (adjustStatus_ == ADJUST_SUCCESS ? "Converged" : "Failed to converge after maximum iterations") << "\n";
        char buf[64];
        snprintf(buf, sizeof(buf), "00:00:%09.6f", total_ms_ / 1e3);
        var("Total time") << buf << "\n\n";
        var("Number of unknown parameters") << stats_.unknown_params << "\n";
        var("Number of measurements") << stats_.measurement_params << "  (" << stats_.outliers << " potential outliers)\n";
        var("Degrees of freedom") << stats_.dof << "\n";
        var("Chi squared") << std::fixed << std::setprecision(2) << stats_.chi_squared << "\n";
        var("Rigorous Sigma Zero") << std::fixed << std::setprecision(3) << stats_.sigma_zero << "\n";
        var("Global (Pelzer) Reliability") << std::fixed << std::setprecision(3) << stats_.global_pelzer
                                           << "   (excludes non redundant measurements)\n\n";
        if (a_.adjust_mode == Phased_Block_1Mode) {   // no global test in block-1 mode (ADJ:7140-7147)
            os << "\n";
            return;
        }
        std::ostringstream t;
        t << "Chi-Square test (" << std::fixed << std::setprecision(1) << a_.confidence_interval << "%)";
        var(t.str().c_str()) << std::fixed << std::setprecision(3) << chiLower_ << " < " << stats_.sigma_zero << " < " << chiUpper_
                             << "          *** " << (passFail_ == 0 ? "PASSED" : (passFail_ == 1 ? "WARNING" : "FAILED")) << " ***\n\n";
    }

    // PrintAdjustedNetworkMeasurements (PRN:494-533): every measurement; block-1 mode reports the measurements of the
    // first block only; --output-msr-blocks prints one table per .seg block (the block's CML)
    void PrintAdjustedNetworkMeasurements(std::ostream& os) const
    {
        const bool phased = a_.adjust_mode != SimultaneousMode && !seg_.cml.empty();
        if (!phased || (!a_.output_msr_blocks && a_.adjust_mode != Phased_Block_1Mode)) {
            PrintAdjMeasurements(os, nullptr, -1);
            return;
        }
        // block of every record: a measurement spans the records from its first one (listed in a block's CML) up to the
        // next listed first record
        std::vector<int32_t> rec_block(msr_.size(), -1);
        for (size_t b = 0; b < seg_.cml.size(); ++b)
            for (uint32_t f : seg_.cml[b])
                if (f < rec_block.size())
                    rec_block[f] = (int32_t)b;
        for (size_t i = 0, cur = (size_t)-1; i < rec_block.size(); ++i) {
            if (rec_block[i] >= 0)
                cur = (size_t)rec_block[i];
            else if (cur != (size_t)-1)
                rec_block[i] = (int32_t)cur;
        }
        for (size_t b = 0; b < seg_.cml.size(); ++b) {
            if (a_.output_msr_blocks)
                os << "\nBlock " << b + 1 << "\n";
            PrintAdjMeasurements(os, &rec_block, (int32_t)b);
            if (a_.adjust_mode == Phased_Block_1Mode)
                break;
        }
    }

    void PrintAdjMeasurements(std::ostream& os, const std::vector<int32_t>* rec_block, int32_t block) const
    {
        os << "\nAdjusted Measurements\n------------------------------------------\n\n";
        char buf[512];
        snprintf(buf, sizeof(buf), "%-2s%-20s%-20s%-20s%-3s%-2s%19s%19s%12s%13s%13s%13s%11s%12s%14s%7s", "M", "Station 1", "Station 2",
                 "Station 3", "*", "C", "Measured", "Adjusted", "Correction", "Meas. SD", "Adj. SD", "Corr. SD", "N-stat", "Pelzer Rel",
                 "Pre Adj Corr", "Out");
        os << buf << "\n" << std::string(200, '-') << "\n";
        const double crit = stats_.critical_value;
        for (size_t i = 0; i < msr_.size(); ++i) {
            const dna_msr_t& m = msr_[i];
            if (m.ignore || m.measStart > 2)   // covariance records of X / Y clusters carry no row
                continue;
            if (rec_block && (*rec_block)[i] != block)
                continue;
            const char t = m.measType;
            const bool gnss = t == 'G' || t == 'X' || t == 'Y';
            if (t == 'Y' && m.measStart == 0 && (m.station3 == DNA_LLH_TYPE || m.station3 == DNA_LLh_TYPE) && i + 2 < msr_.size()) {
                PrintAdjMeasurements_YLLH(os, i, crit);
                i += 2;
                continue;
            }
            // angles (A B D K V Z) and astronomic / geodetic latitudes and longitudes (I J P Q) print as d m s, their
            // corrections and standard deviations in seconds (PrintAdjMeasurementsAngular, PRN:195-201, 2302-2350)
            const bool angular = std::strchr("ABDKVZIJPQ", t) != nullptr;
            char comp = gnss ? "XYZ"[(int)m.measStart] : ' ';
            const double var = !gnss ? m.term2 : (m.measStart == 0 ? m.term2 : (m.measStart == 1 ? m.term3 : m.term4));
            const char* s1 = stn_[m.station1].stationName;
            const char* s2 = (m.measurementStations >= 2 && t != 'Y') ? stn_[m.station2].stationName : "";
            const char* s3 = (m.measurementStations >= 3 && t == 'A') ? stn_[m.station3].stationName : "";
            PrintMsrRow(os, t, s1, s2, s3, comp, angular, m.preAdjMeas, m.measAdj, m.measCorr, var, m.measAdjPrec, m.residualPrec, m.NStat,
                        m.PelzerRel, m.preAdjCorr, crit);
        }
        os << "\n";
    }

    void PrintMsrRow(std::ostream& os, char t, const char* s1, const char* s2, const char* s3, char comp, bool angular, double measured,
                     double adjusted, double corr, double var, double adj_prec, double res_prec, double nstat, double pelzer,
                     double pre_adj_corr, double crit) const
    {
        const double SEC = 3.14159265358979323846 / 180.0 / 3600.0, unit = angular ? SEC : 1.0;
        char meas[32], adjd[32], buf[512];
        if (angular) {
            snprintf(meas, sizeof(meas), "%s", dms_spaced(measured).c_str());
            snprintf(adjd, sizeof(adjd), "%s", dms_spaced(adjusted).c_str());
        } else {
            snprintf(meas, sizeof(meas), "%.4f", measured);
            snprintf(adjd, sizeof(adjd), "%.4f", adjusted);
        }
        snprintf(buf, sizeof(buf), "%-2c%-20s%-20s%-20s%-3s%-2c%19s%19s%12.4f%13.4f%13.4f%13.4f%11.2f%12.2f%14.4f%7s", t, s1, s2, s3, "", comp, meas,
                 adjd, corr / unit, std::sqrt(var) / unit, std::sqrt(std::fabs(adj_prec)) / unit, std::sqrt(res_prec) / unit, nstat, pelzer,
                 pre_adj_corr / unit, std::fabs(nstat) > crit ? "*" : "");
        os << buf << "\n";
    }

    // A point of a Y cluster that was supplied as latitude / longitude / height is reported in that form
    // (PrintAdjMeasurements_YLLH PRN:2488-2660, ReduceYLLHMeasurementsforPrinting ADJ:9981-10046): the adjusted Cartesian
    // point goes back to geographic (orthometric height for LLH: minus the geoid separation), the corrections are taken
    // against the original values kept in preAdjMeas, and the variances of the measurement (its 3x3 Cartesian block) and
    // of the adjusted measurement (its three Cartesian variances) are propagated to geographic with the Jacobian at the
    // adjusted position; N-stat and Pelzer reliability are then recomputed in that frame.
    void PrintAdjMeasurements_YLLH(std::ostream& os, size_t i, double crit) const
    {
        const dna_msr_t* r = &msr_[i];
        const dna_stn_t& st = stn_[r->station1];
        gadj_opts o;
        gadj_default_opts(&o);
        const gadj::Ellipsoid ell = gadj::make_ellipsoid(o.semi_major, o.inv_flattening);
        double llh[3];
        gadj::cart_to_geo(ell, r[0].measAdj, r[1].measAdj, r[2].measAdj, llh);
        // d(XYZ)/d(lat, lon, h) at the adjusted position (FormCarttoGeoRotationMatrix, MFN:204-233) and its inverse
        const double lat = llh[0], lon = llh[1], h = llh[2];
        const double sl = std::sin(lat), cl = std::cos(lat), so = std::sin(lon), co = std::cos(lon);
        const double nu = gadj::prime_vertical(ell, lat), ome = 1.0 - ell.e2;
        const double t1b = ell.a * ell.e2 * sl * cl, t1c = std::pow(1.0 - ell.e2 * sl * sl, 1.5);
        const double J[9] = {t1b * cl * co / t1c - (nu + h) * sl * co, -(nu + h) * cl * so, cl * co,
                             t1b * cl * so / t1c - (nu + h) * sl * so, (nu + h) * cl * co,  cl * so,
                             t1b * ome * sl / t1c + (nu * ome + h) * cl, 0.0,               sl};
        const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
        const double Ji[9] = {(J[4] * J[8] - J[5] * J[7]) / det, (J[2] * J[7] - J[1] * J[8]) / det, (J[1] * J[5] - J[2] * J[4]) / det,
                              (J[5] * J[6] - J[3] * J[8]) / det, (J[0] * J[8] - J[2] * J[6]) / det, (J[2] * J[3] - J[0] * J[5]) / det,
                              (J[3] * J[7] - J[4] * J[6]) / det, (J[1] * J[6] - J[0] * J[7]) / det, (J[0] * J[4] - J[1] * J[3]) / det};
        auto to_geo_diag = [&](const double* V, double* out) {     // diag(Ji V Ji^T)
            for (int a = 0; a < 3; ++a) {
                double s = 0.0;
                for (int x = 0; x < 3; ++x)
                    for (int y = 0; y < 3; ++y)
                        s += Ji[3 * a + x] * V[3 * x + y] * Ji[3 * a + y];
                out[a] = s;
            }
        };
        const double Vm[9] = {r[0].term2, r[1].term2, r[2].term2, r[1].term2, r[1].term3, r[2].term3, r[2].term2, r[2].term3, r[2].term4};
        const double Va[9] = {r[0].measAdjPrec, 0, 0, 0, r[1].measAdjPrec, 0, 0, 0, r[2].measAdjPrec};
        double var[3], adjp[3];
        to_geo_diag(Vm, var);
        to_geo_diag(Va, adjp);
        double adj[3] = {lat, lon, h};
        const bool ortho = r->station3 == DNA_LLH_TYPE;
        if (ortho && std::fabs((double)st.geoidSep) > 1.0e-4)
            adj[2] -= st.geoidSep;
        const char comp[3] = {'P', 'L', ortho ? 'H' : 'h'};
        for (int q = 0; q < 3; ++q) {
            const double corr = adj[q] - r[q].preAdjMeas;
            const double resp = std::fabs(var[q] - adjp[q]);
            double pelzer = std::sqrt(var[q]) / std::sqrt(resp);
            if (!(pelzer >= 0.0) || pelzer > 700.0)
                pelzer = 999.99;
            PrintMsrRow(os, 'Y', st.stationName, "", "", comp[q], q < 2, r[q].preAdjMeas, adj[q], corr, var[q], adjp[q], resp, corr / std::sqrt(resp),
                        pelzer, r[q].preAdjCorr, crit);
        }
    }

    // reference-frame name for file names: GDA2020 / GDA94 from the EPSG code of the station file, else "EPSG<code>"
    std::string frame_name() const
    {
        const std::string e = bst_meta_.epsgCode;
        if (e == "7843")
            return "GDA2020";
        if (e == "4283" || e == "4939")
            return "GDA94";
        return e.empty() ? "GDA2020" : "EPSG" + e;
    }
    // YY:DDD:SSSSS of a dd.mm.yyyy date (DateSINEXFormat, dnachronutils.hpp:98-123); today with seconds when `today`
    static std::string sinex_date(const std::string& ddmmyyyy, bool today)
    {
        int d = 1, m = 1, y = 2020;
        long sec = 0;
        if (today) {
            const std::time_t t = std::time(nullptr);
            std::tm g{};
            gmtime_r(&t, &g);
            d = g.tm_mday;
            m = g.tm_mon + 1;
            y = g.tm_year + 1900;
            sec = g.tm_hour * 3600L + g.tm_min * 60L + g.tm_sec;
        } else if (sscanf(ddmmyyyy.c_str(), "%d.%d.%d", &d, &m, &y) != 3) {
            d = m = 1;
            y = 2020;
        }
        static const int cum[2][12] = {{0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334}, {0, 31, 60, 91, 121, 152, 182, 213, 244, 274, 305, 335}};
        const int leap = (y % 400 == 0 || (y % 100 != 0 && y % 4 == 0)) ? 1 : 0;
        char b[32];
        snprintf(b, sizeof(b), "%02d:%03d:%05ld", y % 100, cum[leap][(m - 1) % 12] + d, sec);
        return b;
    }
    // FormatDmsString(RadtoDms(x), 5, spaces): "ddd mm ss.s"
    static std::string dms_spaced5(double rad)
    {
        const double deg = std::fabs(rad) * 180.0 / 3.14159265358979323846;
        const long long units = std::llround(deg * 3600.0 * 10.0);
        const long long d = units / 36000, rem = units % 36000, mi = rem / 600, s10 = rem % 600;
        char b[48];
        snprintf(b, sizeof(b), "%s%lld %02lld %02lld.%lld", rad < 0 ? "-" : "", d, mi, s10 / 10, s10 % 10);
        return b;
    }

    // "ddd mm ss.ssss" (FormatDmsString with spaces on a RadtoDms value, 4 decimals of a second)
    static std::string dms_spaced(double rad)
    {
        const double deg = std::fabs(rad) * 180.0 / 3.14159265358979323846;
        long long units = std::llround(deg * 3600.0 * 10000.0);   // ten-thousandths of a second: carries are exact
        const long long d = units / (3600LL * 10000), rem = units % (3600LL * 10000);
        const long long mi = rem / (60LL * 10000), sec = rem % (60LL * 10000);
        char b[48];
        snprintf(b, sizeof(b), "%s%lld %02lld %02lld.%04lld", rad < 0 ? "-" : "", d, mi, sec / 10000, sec % 10000);
        return b;
    }

    void PrintAdjStations(std::ostream& os, const std::vector<uint32_t>* subset) const
    {   // PrintAdjStation (PRN:3917-4070): PLHhXYZ + SD(e,n,up) = sqrt diag(R^T Q R), geoid uncertainty added to up
        os << "\nAdjusted Coordinates\n------------------------------------------\n\n";
        char buf[512];
        snprintf(buf, sizeof(buf), "%-20s%-5s%14s%15s%11s%11s%15s%15s%15s%12s%10s%10s  %s", "Station", "Const", "Latitude", "Longitude",
                 "H(Ortho)", "h(Ellipse)", "X", "Y", "Z", "SD(e)", "SD(n)", "SD(up)", "Description");
        os << buf << "\n" << std::string(211, '-') << "\n";
        const size_t count = subset ? subset->size() : stn_.size();
        for (size_t n = 0; n < count; ++n) {
            const size_t i = subset ? (*subset)[n] : n;
            const dna_stn_t& s = stn_[i];
            const double* q = &vcv_[9 * i];
            double lat = s.currentLatitude, lon = s.currentLongitude, h = s.currentHeight;
            double sl = std::sin(lat), cl = std::cos(lat), so = std::sin(lon), co = std::cos(lon);
            double R[3][3] = {{-so, -sl * co, cl * co}, {co, -sl * so, cl * so}, {0, cl, sl}};  // local -> cart
            double sd[3];
            for (int k = 0; k < 3; ++k) {
                double v = 0;
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        v += R[a][k] * q[3 * a + b] * R[b][k];
                if (k == 2)
                    v += (double)s.geoidSepUnc * s.geoidSepUnc;
                sd[k] = std::sqrt(std::fabs(v));
            }
            char cst[4] = {s.stationConst[0], s.stationConst[1], s.stationConst[2], 0};
            snprintf(buf, sizeof(buf), "%-20s%-5s%14s%15s%11.4f%11.4f%15.4f%15.4f%15.4f%12.4f%10.4f%10.4f  %s", s.stationName, cst,
                     hp_dms(lat).c_str(), hp_dms(lon).c_str(), h - (double)s.geoidSep, h, est_[3 * i], est_[3 * i + 1], est_[3 * i + 2],
                     sd[0], sd[1], sd[2], s.description);
            os << buf << "\n";
        }
        os << "\n";
    }

    adjust_settings a_;
    gadj_ctx* ctx_ = nullptr;
    gadj_info info_{};
    gadj_stats stats_{};
    std::vector<dna_stn_t> stn_;
    std::vector<dna_msr_t> msr_;
    dnafiles::BinaryMeta bst_meta_, bms_meta_;
    dnafiles::Segmentation seg_;
    std::string bst_file_, bms_file_;
    std::vector<gadj_iter_result> iterations_;
    std::vector<double> est_, vcv_, apriori_llh_, apriori_xyz_;
    double maxCorr_ = 0, total_ms_ = 0, chiUpper_ = 0, chiLower_ = 0;
    int passFail_ = 0;
    ADJUST_STATUS adjustStatus_ = ADJUST_SUCCESS;
};

}  // namespace dynadjust_b200
