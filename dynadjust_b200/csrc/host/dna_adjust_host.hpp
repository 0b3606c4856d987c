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
#include <array>
#include <charconv>
#include <map>
#include <tuple>
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
    bool stage = false, multi_thread = false;    // --staged-adjustment / --multi-thread: phased; they name the outputs (WRAP:687-704)
    std::string bst_file, bms_file, seg_file;    // --binary-stn-file / --binary-msr-file / --seg-file: override the network name
    std::string stage_path;                      // --stage-path: where <net>-rva.mtx / <net>-pam.mtx go (default: the output folder)
    bool report_results = false;                 // --report-results: print the last adjustment again (WRAP:607-614)
    double iteration_threshold = (double)0.0005f;   // float in the reference (dnaoptions.hpp:432)
    uint32_t max_iterations = 10;
    double free_std_dev = 10.0, fixed_std_dev = 1.0e-6, confidence_interval = 95.0;
    bool scale_normals_to_unity = false;
    bool output_adj_msr = false;
    bool output_stn_blocks = false;        // --output-stn-blocks: phased modes print the station tables block by block
    bool output_msr_blocks = false;        // --output-msr-blocks: likewise the adjusted measurements
    bool output_pos_uncertainty = false;   // --output-pos-uncertainty: <net>.<mode>.apu
    bool output_corrections = false;       // --output-corrections-file: <net>.<mode>.cor
    bool export_xml_stn = false, export_xml_msr = false, export_dna_stn = false, export_dna_msr = false;   // --export-xml-stn-file ... (WRAP:377-447)
    bool output_json = false;              // --output-json: JSONL siblings of the text reports
    bool export_sinex = false;             // --export-sinex-file: <net>[-block<k>].<frame>.snx with the dense block variance matrix
    bool apu_vcv_enu = false;              // --output-apu-vcv-units ENU (default XYZ)
    bool output_pu_covariances = false;    // --output-all-covariances: covariance blocks between the stations of a block in the .apu
    double hz_corr_threshold = 0.0, vt_corr_threshold = 0.0;   // dnaoptions.hpp:510
    bool update_binary_files = true;
    bool update_project_file = true;       // <net>.dnaproj is rewritten with the settings of the run (WRAP:1456-1466)
    std::string type_b_global, type_b_file; // --type-b-sd-global "e,n,up" (metres, 1 sigma), --type-b-sd-file <file> (dnaoptions-interface.hpp)
    std::string station_constraints;       // --constraints "STN1,CCC,STN2,FFC" (dnaoptions.hpp:481)
    // report layout (output_settings, dnaoptions.hpp:496-516)
    int sort_adj_msr = 0;                  // --sort-adj-msr-field 0 file order | 1 type | 2 inst | 3 targ | 4 value | 5 correction | 6 adj sd | 7 n-stat
    int adj_gnss_units = 0;                // --output-adj-gnss-units 0 XYZ | 1 ENU | 2 AED | 3 ADU
    bool adj_msr_tstat = false;            // --output-tstat-adj-msr
    bool output_msr_to_stn = false;        // --output-msr-to-stn
    int sort_msr_to_stn = 0;               // --sort-msr-to-stn-field 0 file order | 1 name | 2 count | 3 count descending
    bool stn_corrections = false;          // --stn-corrections: Corr(e) Corr(n) Corr(up) columns in the station tables
    std::string stn_coord_types = "PLHhXYZ";   // --stn-coord-types
    bool sort_stn_orig_order = false;      // --sort-stn-orig-order
    int angular_type_stn = 0, angular_type_msr = 0, dms_format_msr = 0;   // 0 dms | 1 decimal degrees; 0 "d m s" | 1 symbols | 2 d.mmsss
    int precision_seconds_stn = 5, precision_metres_stn = 4, precision_seconds_msr = 4, precision_metres_msr = 4;
    bool iter_adj_stn = false, iter_adj_stat = false, iter_adj_msr = false, iter_cmp_msr = false;   // --output-iter-*
    bool output_ignored_msrs = false;      // --output-ignored-msrs
    bool database_ids = false;             // --output-database-ids: measurement / cluster ids of <net>.dbid beside every row
    std::string comments;                  // --comments
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
        auto in_folder = [&](const std::string& f) { return f.find('/') == std::string::npos ? a_.input_folder + "/" + f : f; };
        bst_file_ = a_.bst_file.empty() ? base + ".bst" : in_folder(a_.bst_file);
        bms_file_ = a_.bms_file.empty() ? base + ".bms" : in_folder(a_.bms_file);
        dnafiles::load_binary(bst_file_, stn_, bst_meta_);
        dnafiles::load_binary(bms_file_, msr_, bms_meta_);
        ApplyConstraints();
        if (a_.database_ids)
            LoadDatabaseId();
        gadj_opts o;
        gadj_default_opts(&o);
        o.fixed_std_dev = a_.fixed_std_dev;
        o.free_std_dev = a_.free_std_dev;
        o.iteration_threshold = a_.iteration_threshold;
        o.max_iterations = a_.max_iterations;
        o.confidence_interval = a_.confidence_interval;
        o.scale_normals_to_unity = 1;   // internal equilibration is always safe; the flag is accepted for compatibility
        if ((a_.export_sinex || a_.export_xml_msr || a_.export_dna_msr || a_.output_pu_covariances) && a_.adjust_mode == SimultaneousMode) {
            // the SINEX / Y cluster files of a simultaneous adjustment carry the full variance matrix (PRN:2944-2946, 3085-3087): one dense front
            if (stn_.size() > 12000)
                SignalExceptionAdjustment("--export-sinex-file / --export-*-msr-file / --output-all-covariances in simultaneous mode need the full variance matrix of the network "
                                          "(dense); segment the network (dnasegment) and run --phased-adjustment for per-block files");
            o.ordering = GADJ_ORDER_DENSE;
        }
        if (gadj_create(&o, &ctx_))
            SignalExceptionAdjustment(gadj_last_error(nullptr));
        check(gadj_set_stations(ctx_, stn_.data(), (uint32_t)stn_.size()));
        check(gadj_set_measurements(ctx_, msr_.data(), msr_.size()));
        check(gadj_set_measurements_reduced(ctx_, bms_meta_.reduced ? 1 : 0));   // isFirstTimeAdjustment_ (ADJ:296)
        if (a_.adjust_mode != SimultaneousMode) {
            dnafiles::load_seg(a_.seg_file.empty() ? base + ".seg" : in_folder(a_.seg_file), seg_);
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
        iter_pre_.clear();
        iter_post_.clear();
        const bool iter_reports = a_.iter_adj_stn || a_.iter_adj_stat || a_.iter_adj_msr || a_.iter_cmp_msr;
        for (uint32_t i = 0; i < a_.max_iterations; ++i) {
            std::ostringstream pre, post;
            if (a_.iter_cmp_msr) {
                // computed measurements at the start of the iteration (ADJ:2443-2445): the a-priori evaluation before the
                // first solve, afterwards the records re-linearised by the statistics of the previous iteration
                if (i == 0)
                    check(gadj_compute_measurements(ctx_));
                PrintMsrTableHeader(pre, "Computed Measurements (a-priori)", 1);
                PrintMeasurementRecords(pre, CollectMeasurements(nullptr, -1, false), 1);
                pre << "\n";
            }
            auto ti = std::chrono::steady_clock::now();
            gadj_iter_result r;
            check(gadj_iterate(ctx_, i == 0 ? GADJ_ITER_NORMALS : 0, &r));
            r.ms_inverse = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - ti).count();  // wall
            iterations_.push_back(r);
            maxCorr_ = r.max_corr;
            UpdateIterationDiagnostics();
            iter_pre_.push_back(pre.str());
            iter_post_.push_back(std::string());
            if (std::fabs(r.max_corr) <= a_.iteration_threshold)
                break;
            if (iter_reports && i + 1 < a_.max_iterations) {
                // --output-iter-adj-stat / -msr / -stn: statistics, adjusted measurements and stations of an iteration that
                // is followed by another (ADJ:2483-2502); needs the rigorous variances of this iteration
                check(gadj_form_inverse(ctx_));
                GenerateStatistics();
                if (a_.iter_adj_stat)
                    PrintStatisticsSummary(post, false);
                if (a_.iter_adj_msr)
                    PrintAdjMeasurements(post, nullptr, -1);
                if (a_.iter_adj_stn)
                    PrintAdjStations(post, nullptr);
                iter_post_.back() = post.str();
            }
        }
        if (iterations_.size() == a_.max_iterations && std::fabs(maxCorr_) > a_.iteration_threshold)
            adjustStatus_ = ADJUST_MAX_ITERATIONS_EXCEEDED;   // ADJ:2523-2525
        if (adjustStatus_ == ADJUST_SUCCESS)
            check(gadj_form_inverse(ctx_));                    // rigorous variances (v_rigorousVariances_)
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
        raw_vcv_ = vcv_;                              // before type B uncertainties: adjusted-measurement precisions use these
        ApplyTypeBUncertainties();
        ComputeTestStat();
        if (a_.adj_msr_tstat) {                       // Student's t = n-stat / sqrt(sigma zero) (UpdateMsrTstatistic ADJ:6914-7092)
            const double sz = std::sqrt(stats_.sigma_zero);
            for (dna_msr_t& m : msr_)
                if (!m.ignore && m.measStart <= 2)
                    m.TStat = std::fabs(sz) < 1.0e-10 ? 0.0 : m.NStat / sz;
        }
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
        case PhasedMode: return a_.stage ? "phased-stage" : (a_.multi_thread ? "phased-mt" : "phased");
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
        if (report_mode_)
            adj << "\n+ Loading network files\n+ Printing results of the last adjustment\n\n";
        else {
            adj << "\n+ Initialising adjustment\n+ Loading network files\n+ Allocating memory\n\n+ Preparing for adjustment...  done.\n";
            adj << "+ Commencing " << (a_.adjust_mode == SimultaneousMode ? "simultaneous" : "phased") << " adjustment\n\n";
        }
        if (a_.adj_gnss_units != 0 && a_.output_adj_msr)
            ComputeBaselinePrecisions();
        if (report_mode_ && (a_.export_sinex || a_.export_xml_msr || a_.export_dna_msr || a_.output_pu_covariances))
            SignalExceptionAdjustment("Report results: the block variance matrices (--export-sinex-file, --export-*-msr-file, --output-all-covariances) "
                                      "are formed by an adjustment only; run the adjustment with these options.");
        for (size_t i = 0; i < iterations_.size(); ++i) {
            PrintIteration(adj, (uint32_t)i + 1, iterations_[i]);
            adj << iter_pre_[i] << iter_post_[i];
        }
        PrintStatistics(adj);
        if (a_.output_adj_msr)
            PrintAdjustedNetworkMeasurements(adj);
        if (a_.output_adj_msr && a_.output_ignored_msrs)
            PrintIgnoredAdjMeasurements(adj);
        if (a_.output_msr_to_stn)
            PrintMeasurementsToStation(adj);
        std::ofstream xyz(stem + ".xyz");
        PrintOutputFileHeaderInfo(xyz, "DYNADJUST COORDINATE OUTPUT FILE", stem + ".xyz");
        PrintAdjustedNetworkStations(adj, xyz);
        if (a_.output_pos_uncertainty)
            PrintPositionalUncertainty(stem + ".apu");
        if (a_.output_corrections)
            PrintNetworkStationCorrections(stem + ".cor");
        if (a_.output_json)
            PrintJsonReports(stem);
        if (a_.export_xml_stn)
            PrintEstimatedStationCoordinatestoDNAXML(stem + ".adj.stn.xml", true, stem + ".adj");
        if (a_.export_xml_msr)
            PrintEstimatedStationCoordinatestoDNAXML_Y(stem + ".adj.msr.xml", true, stem + ".adj");
        if (a_.export_dna_stn)
            PrintEstimatedStationCoordinatestoDNAXML(stem + ".adj.stn", false, stem + ".adj");
        if (a_.export_dna_msr)
            PrintEstimatedStationCoordinatestoDNAXML_Y(stem + ".adj.msr", false, stem + ".adj");
        if (a_.export_sinex)
            PrintEstimatedStationCoordinatestoSNX();
    }

    // ---- precision of the adjusted G / X baselines, full 3x3 (v_precAdjMsrsFull_; Precision_Adjusted_GNSS_bsl MFN:255-297):
    // Q11 + Q22 - Q12 - Q21 from the station and pair blocks of the rigorous variances, fetched in one bulk call
    void ComputeBaselinePrecisions()
    {
        if (!pam_rec_.empty() || !ctx_)
            return;
        std::vector<uint32_t> si, sj;
        for (size_t i = 0; i < msr_.size();) {
            const size_t span = MeasurementSpan(i);
            const dna_msr_t& m = msr_[i];
            if (!m.ignore && (m.measType == 'G' || m.measType == 'X'))
                for (size_t j = i; j + 2 < i + span; j += 3 + 3 * (size_t)msr_[j].vectorCount2) {
                    pam_rec_.push_back((uint32_t)j);
                    si.push_back(msr_[j].station1);
                    sj.push_back(msr_[j].station2);
                }
            i += span;
        }
        std::vector<double> q12(9 * si.size());
        if (!si.empty())
            check(gadj_get_pair_vcvs(ctx_, si.size(), si.data(), sj.data(), q12.data()));
        pam_.resize(6 * si.size());
        static const int ua[6] = {0, 0, 0, 1, 1, 2}, ub[6] = {0, 1, 2, 1, 2, 2};
        for (size_t p = 0; p < si.size(); ++p)
            for (int t = 0; t < 6; ++t) {
                const int a = ua[t], b = ub[t];
                pam_[6 * p + t] = raw_vcv_[9 * (size_t)si[p] + 3 * a + b] + raw_vcv_[9 * (size_t)sj[p] + 3 * a + b] - q12[9 * p + 3 * a + b] -
                                  q12[9 * p + 3 * b + a];
            }
    }
    void BaselinePrecision(size_t rec, double* Va) const
    {
        auto it = std::lower_bound(pam_rec_.begin(), pam_rec_.end(), (uint32_t)rec);
        if (it == pam_rec_.end() || *it != rec)
            throw std::runtime_error("the precision of an adjusted baseline is not available (re-run the adjustment)");
        const double* v = &pam_[6 * (size_t)(it - pam_rec_.begin())];
        const double M[9] = {v[0], v[1], v[2], v[1], v[3], v[4], v[2], v[4], v[5]};
        std::memcpy(Va, M, sizeof(M));
    }

    // ---- <net>-rva.mtx / <net>-pam.mtx (SerialiseAdjustedVarianceMatrices ADJ:6770-6799): what --report-results needs to
    // print the last adjustment again without solving — the solution summary and the 3x3 variance block of every station
    // (rva), the precisions of the adjusted baselines (pam).  The reference keeps its dense per-block matrices in these
    // files; they are private to dnaadjust, so the layout here is this program's own: an 8-byte tag, counts, raw doubles.
    struct ReportHeader {
        char tag[8];
        uint64_t nstn, nmsr;
        gadj_stats stats;
        double chi_lower, chi_upper, max_corr, total_ms;
        int32_t pass_fail, status, iterations, mode;
    };
    std::string StagePath(const char* what) const
    {
        return (a_.stage_path.empty() ? a_.output_folder : a_.stage_path) + "/" + a_.network_name + "-" + what + ".mtx";
    }
    void SerialiseAdjustedVarianceMatrices()
    {
        ComputeBaselinePrecisions();
        ReportHeader h{};
        std::memcpy(h.tag, "GADJRVA1", 8);
        h.nstn = stn_.size();
        h.nmsr = msr_.size();
        h.stats = stats_;
        h.chi_lower = chiLower_, h.chi_upper = chiUpper_, h.max_corr = maxCorr_, h.total_ms = total_ms_;
        h.pass_fail = passFail_, h.status = (int32_t)adjustStatus_, h.iterations = (int32_t)iterations_.size(), h.mode = a_.adjust_mode;
        std::ofstream rva(StagePath("rva"), std::ios::binary);
        rva.write(reinterpret_cast<const char*>(&h), sizeof(h));
        rva.write(reinterpret_cast<const char*>(raw_vcv_.data()), (std::streamsize)(raw_vcv_.size() * sizeof(double)));
        std::memcpy(h.tag, "GADJPAM1", 8);
        h.nstn = pam_rec_.size();
        std::ofstream pam(StagePath("pam"), std::ios::binary);
        pam.write(reinterpret_cast<const char*>(&h), sizeof(h));
        pam.write(reinterpret_cast<const char*>(pam_rec_.data()), (std::streamsize)(pam_rec_.size() * sizeof(uint32_t)));
        pam.write(reinterpret_cast<const char*>(pam_.data()), (std::streamsize)(pam_.size() * sizeof(double)));
        if (!rva || !pam)
            SignalExceptionAdjustment("SerialiseAdjustedVarianceMatrices(): could not write " + StagePath("rva") + " / " + StagePath("pam"));
    }

    // ---- --report-results (WRAP:607-614, 1382-1384; DeSerialiseAdjustedVarianceMatrices ADJ:6720-6767): the binary files of
    // the last adjustment already hold the adjusted coordinates and the measurement statistics; with the two .mtx files
    // every report is printed again.  No solve, no device.
    void LoadLastAdjustment(const adjust_settings& s)
    {
        a_ = s;
        report_mode_ = true;
        const std::string base = a_.input_folder + "/" + a_.network_name;
        auto in_folder = [&](const std::string& f) { return f.find('/') == std::string::npos ? a_.input_folder + "/" + f : f; };
        bst_file_ = a_.bst_file.empty() ? base + ".bst" : in_folder(a_.bst_file);
        bms_file_ = a_.bms_file.empty() ? base + ".bms" : in_folder(a_.bms_file);
        dnafiles::load_binary(bst_file_, stn_, bst_meta_);
        dnafiles::load_binary(bms_file_, msr_, bms_meta_);
        if (a_.adjust_mode != SimultaneousMode)
            dnafiles::load_seg(a_.seg_file.empty() ? base + ".seg" : in_folder(a_.seg_file), seg_);
        if (a_.database_ids)
            LoadDatabaseId();
        ReportHeader h{};
        std::ifstream rva(StagePath("rva"), std::ios::binary);
        if (!rva || !rva.read(reinterpret_cast<char*>(&h), sizeof(h)) || std::memcmp(h.tag, "GADJRVA1", 8) != 0)
            SignalExceptionAdjustment("Report results: " + StagePath("rva") + " was not found or is not a variance file of this program.\n"
                                      "  Run an adjustment first.");
        if (h.nstn != stn_.size() || h.nmsr != msr_.size())
            SignalExceptionAdjustment("Report results: " + StagePath("rva") + " does not belong to the binary station and measurement files.");
        raw_vcv_.resize(9 * stn_.size());
        rva.read(reinterpret_cast<char*>(raw_vcv_.data()), (std::streamsize)(raw_vcv_.size() * sizeof(double)));
        if (!rva)
            SignalExceptionAdjustment("Report results: " + StagePath("rva") + " is truncated.");
        stats_ = h.stats;
        chiLower_ = h.chi_lower, chiUpper_ = h.chi_upper, maxCorr_ = h.max_corr, total_ms_ = h.total_ms;
        passFail_ = h.pass_fail, adjustStatus_ = (ADJUST_STATUS)h.status, last_iterations_ = (uint32_t)h.iterations;
        ReportHeader hp{};
        std::ifstream pam(StagePath("pam"), std::ios::binary);
        if (pam && pam.read(reinterpret_cast<char*>(&hp), sizeof(hp)) && std::memcmp(hp.tag, "GADJPAM1", 8) == 0 && hp.nmsr == msr_.size()) {
            pam_rec_.resize(hp.nstn);
            pam_.resize(6 * hp.nstn);
            pam.read(reinterpret_cast<char*>(pam_rec_.data()), (std::streamsize)(pam_rec_.size() * sizeof(uint32_t)));
            pam.read(reinterpret_cast<char*>(pam_.data()), (std::streamsize)(pam_.size() * sizeof(double)));
            if (!pam)
                pam_rec_.clear(), pam_.clear();
        }
        const gadj::Ellipsoid ell = Ellipsoid();
        est_.resize(3 * stn_.size());
        apriori_llh_.resize(3 * stn_.size());
        for (size_t i = 0; i < stn_.size(); ++i) {
            gadj::geo_to_cart(ell, stn_[i].currentLatitude, stn_[i].currentLongitude, stn_[i].currentHeight, &est_[3 * i]);
            apriori_llh_[3 * i] = stn_[i].currentLatitude;
            apriori_llh_[3 * i + 1] = stn_[i].currentLongitude;
            apriori_llh_[3 * i + 2] = stn_[i].currentHeight;
        }
        apriori_xyz_ = est_;
        vcv_ = raw_vcv_;
        ApplyTypeBUncertainties();
        info_.nstations = (uint32_t)stn_.size();
        info_.nfronts = 0;
    }
    bool ReportMode() const { return report_mode_; }

    // ---- JSONL siblings of the text reports (--output-json; DynAdjustJsonPrinter dnaadjust_json_printer.cpp:40-615): one JSON
    // object per line, keys in alphabetical order and numbers in shortest round-trip form as the reference's JSON library
    // writes them.  <adj>.jsonl: header, DnaStatistics, one DnaMeasurement per measurement of the adjusted-measurements
    // table, one DnaStation per station; <xyz>.jsonl, <apu>.jsonl, <cor>.jsonl: header and one DnaStation per station.
    struct Json {
        enum Kind { Null, Bool, Int, Real, Str, Arr, Obj } kind = Null;
        bool b = false;
        long long i = 0;
        double d = 0.0;
        std::string s;
        std::vector<Json> a;
        std::map<std::string, Json> o;
        Json() = default;
        Json(bool v) : kind(Bool), b(v) {}
        Json(int v) : kind(Int), i(v) {}
        Json(uint32_t v) : kind(Int), i(v) {}
        Json(long long v) : kind(Int), i(v) {}
        Json(int64_t v) : kind(Int), i(v) {}
        Json(double v) : kind(Real), d(v) {}
        Json(const char* v) : kind(Str), s(v) {}
        Json(const std::string& v) : kind(Str), s(v) {}
        static Json array() { Json j; j.kind = Arr; return j; }
        Json& operator[](const char* k) { kind = Obj; return o[k]; }
        void push_back(const Json& v) { kind = Arr; a.push_back(v); }
        void dump(std::string& out) const
        {
            switch (kind) {
            case Null: out += "null"; break;
            case Bool: out += b ? "true" : "false"; break;
            case Int: out += std::to_string(i); break;
            case Real: {
                if (!std::isfinite(d)) {
                    out += "null";
                    break;
                }
                char buf[40];
                auto r = std::to_chars(buf, buf + sizeof(buf), d);
                std::string t(buf, r.ptr);
                if (t.find_first_of(".e") == std::string::npos)
                    t += ".0";
                out += t;
                break;
            }
            case Str:
                out += '"';
                for (char c : s) {
                    if (c == '"' || c == '\\') {
                        out += '\\';
                        out += c;
                    } else if ((unsigned char)c < 0x20) {
                        char e[8];
                        snprintf(e, sizeof(e), "\\u%04x", c);
                        out += e;
                    } else
                        out += c;
                }
                out += '"';
                break;
            case Arr:
                out += '[';
                for (size_t k = 0; k < a.size(); ++k) {
                    if (k)
                        out += ',';
                    a[k].dump(out);
                }
                out += ']';
                break;
            case Obj:
                out += '{';
                {
                    bool firstkey = true;
                    for (const auto& kv : o) {
                        if (!firstkey)
                            out += ',';
                        firstkey = false;
                        Json(kv.first).dump(out);
                        out += ':';
                        kv.second.dump(out);
                    }
                }
                out += '}';
                break;
            }
        }
    };
    static void WriteRecord(std::ostream& os, const char* key, const Json& body)
    {
        Json rec;
        rec[key] = body;
        std::string line;
        rec.dump(line);
        os << line << "\n";
    }
    static std::string Trimmed(const char* p)
    {
        std::string t(p);
        const size_t a = t.find_first_not_of(' '), b = t.find_last_not_of(' ');
        return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
    }
    void JsonHeader(std::ostream& os, const char* report) const
    {
        Json h;
        h["type"] = "Adjustment";
        h["report"] = report;
        h["software"] = "dnaadjust (dynadjust_b200) 1.0";
        h["referenceframe"] = frame_name();
        h["epoch"] = std::string(bst_meta_.epoch);
        WriteRecord(os, "DnaAdjustmentReport", h);
    }
    Json JsonStationIdentity(const dna_stn_t& s) const
    {
        Json j;
        j["Name"] = Trimmed(s.stationName);
        j["Constraints"] = std::string(s.stationConst, strnlen(s.stationConst, 3));
        j["Type"] = "LLH";
        const std::string desc = Trimmed(s.description);
        if (!desc.empty())
            j["Description"] = desc;
        return j;
    }
    static Json Mat3(const double* m)
    {
        Json rows = Json::array();
        for (int r = 0; r < 3; ++r) {
            Json row = Json::array();
            for (int c = 0; c < 3; ++c)
                row.push_back(m[3 * r + c]);
            rows.push_back(row);
        }
        return rows;
    }
    Json JsonUncertainty(size_t i, bool with_geoid) const
    {
        const dna_stn_t& s = stn_[i];
        const double* q = &vcv_[9 * i];
        double ql[9];
        to_local(q, s.currentLatitude, s.currentLongitude, ql);
        if (with_geoid)
            ql[8] += (double)s.geoidSepUnc * s.geoidSepUnc;
        double smaj, smin, az, hz, vt;
        ErrorEllipseParameters(ql, smaj, smin, az);
        PositionalUncertainty(smaj, smin, std::sqrt(std::fabs(ql[8])), hz, vt);
        Json u;
        u["SE"] = std::sqrt(std::fabs(ql[0]));
        u["SN"] = std::sqrt(std::fabs(ql[4]));
        u["SU"] = std::sqrt(std::fabs(ql[8]));
        u["SemiMajor"] = smaj;
        u["SemiMinor"] = smin;
        u["Orientation"] = az;
        u["HzPosU"] = hz;
        u["VtPosU"] = vt;
        u["VarianceLocal"] = Mat3(ql);
        u["VarianceCart"] = Mat3(q);
        return u;
    }
    Json JsonInitial(const dna_stn_t& s) const
    {
        Json j;
        j["Lat"] = rad_to_dms(s.initialLatitude);
        j["Lon"] = rad_to_dms(s.initialLongitude);
        j["Height"] = s.initialHeight;
        return j;
    }
    Json JsonAdjustedStation(size_t i) const
    {
        const dna_stn_t& s = stn_[i];
        Json j = JsonStationIdentity(s), c, adj;
        c["Name"] = Trimmed(s.stationName);
        c["XAxis"] = rad_to_dms(s.currentLatitude);
        c["YAxis"] = rad_to_dms(s.currentLongitude);
        c["Height"] = s.currentHeight;
        j["StationCoord"] = c;
        j["Initial"] = JsonInitial(s);
        adj["X"] = est_[3 * i];
        adj["Y"] = est_[3 * i + 1];
        adj["Z"] = est_[3 * i + 2];
        adj["Lat"] = rad_to_dms(s.currentLatitude);
        adj["Lon"] = rad_to_dms(s.currentLongitude);
        adj["Height"] = s.currentHeight;
        j["Adjusted"] = adj;
        j["Uncertainty"] = JsonUncertainty(i, true);
        return j;
    }
    static bool AngularInput(char t) { return std::strchr("ABDIJKPQVZ", t) != nullptr; }
    void JsonScalarFields(Json& m, const dna_msr_t& r) const
    {
        const double SEC = 3.14159265358979323846 / 180.0 / 3600.0;
        m["Value"] = AngularInput(r.measType) ? rad_to_dms(r.term1) : r.term1;
        m["StdDev"] = AngularInput(r.measType) ? std::sqrt(r.term2) / SEC : std::sqrt(r.term2);
        if (r.ignore)
            m["Ignore"] = true;
        m["Adjusted"] = r.measAdj;
        m["Correction"] = r.measCorr;
        m["AdjustedPrecision"] = r.measAdjPrec;
        m["ResidualPrecision"] = r.residualPrec;
        m["NStat"] = r.NStat;
        m["TStat"] = r.TStat;
        m["PelzerRel"] = r.PelzerRel;
    }
    Json JsonMeasurement(uint32_t first) const
    {
        const dna_msr_t& m0 = msr_[first];
        Json m;
        m["Type"] = std::string(1, m0.measType);
        const std::string oe = Trimmed(std::string(m0.observation_epoch, strnlen(m0.observation_epoch, sizeof(m0.observation_epoch))).c_str());
        if (!oe.empty())
            m["EpochOfObservation"] = oe;
        m["First"] = Trimmed(stn_[m0.station1].stationName);
        if (m0.measType == 'G' || m0.measType == 'X' || m0.measType == 'Y') {
            if (m0.ignore)
                m["Ignore"] = true;
            if (m0.measType != 'Y')
                m["Second"] = Trimmed(stn_[m0.station2].stationName);
            const uint32_t count = std::max<uint32_t>(1u, m0.vectorCount1);
            m["Total"] = count;
            Json comps = Json::array();
            Json trip[6] = {Json::array(), Json::array(), Json::array(), Json::array(), Json::array(), Json::array()};
            size_t j = first;
            for (uint32_t k = 0; k < count && j + 2 < msr_.size(); ++k) {
                const dna_msr_t* r = &msr_[j];
                Json c;
                c["First"] = Trimmed(stn_[r->station1].stationName);
                if (m0.measType != 'Y')
                    c["Second"] = Trimmed(stn_[r->station2].stationName);
                c["X"] = r[0].term1, c["Y"] = r[1].term1, c["Z"] = r[2].term1;
                c["SigmaXX"] = r[0].term2, c["SigmaXY"] = r[1].term2, c["SigmaXZ"] = r[2].term2;
                c["SigmaYY"] = r[1].term3, c["SigmaYZ"] = r[2].term3, c["SigmaZZ"] = r[2].term4;
                if (r->vectorCount2 > 0) {
                    Json covs = Json::array();
                    for (uint32_t q = 0; q < r->vectorCount2; ++q) {
                        const dna_msr_t* cv = r + 3 + 3 * q;
                        Json e;
                        static const char* tag[9] = {"m11", "m12", "m13", "m21", "m22", "m23", "m31", "m32", "m33"};
                        for (int x = 0; x < 3; ++x) {
                            e[tag[3 * x]] = cv[x].term1;
                            e[tag[3 * x + 1]] = cv[x].term2;
                            e[tag[3 * x + 2]] = cv[x].term3;
                        }
                        covs.push_back(e);
                    }
                    c[m0.measType == 'Y' ? "PointCovariance" : "GPSCovariance"] = covs;
                }
                comps.push_back(c);
                const double dna_msr_t::*fld[6] = {&dna_msr_t::measAdj, &dna_msr_t::measCorr, &dna_msr_t::measAdjPrec, &dna_msr_t::NStat, &dna_msr_t::TStat,
                                                   &dna_msr_t::PelzerRel};
                for (int f = 0; f < 6; ++f) {
                    Json t;
                    t["X"] = r[0].*fld[f], t["Y"] = r[1].*fld[f], t["Z"] = r[2].*fld[f];
                    trip[f].push_back(t);
                }
                j += 3 + 3 * (size_t)r->vectorCount2;
            }
            if (m0.measType == 'Y') {
                const std::string coords = Trimmed(std::string(m0.coordType, strnlen(m0.coordType, 4)).c_str());
                if (!coords.empty())
                    m["Coords"] = coords;
                m["Clusterpoint"] = comps;
            } else
                m["GPSBaseline"] = comps;
            static const char* names[6] = {"Adjusted", "Correction", "AdjustedPrecision", "NStat", "TStat", "PelzerRel"};
            for (int f = 0; f < 6; ++f)
                m[names[f]] = trip[f].a.size() == 1 ? trip[f].a[0] : trip[f];
            return m;
        }
        if (m0.measurementStations >= 2)
            m["Second"] = Trimmed(stn_[m0.station2].stationName);
        if (m0.measurementStations >= 3 && m0.measType != 'D')
            m["Third"] = Trimmed(stn_[m0.station3].stationName);
        JsonScalarFields(m, m0);
        if (m0.measType == 'D') {
            Json dirs = Json::array();
            const uint32_t nd = m0.vectorCount1 > 0 ? m0.vectorCount1 - 1 : 0;
            for (uint32_t k = 0; k < nd && first + 1 + k < msr_.size(); ++k) {
                const dna_msr_t& d = msr_[first + 1 + k];
                Json e;
                e["Target"] = Trimmed(stn_[d.station2].stationName);
                JsonScalarFields(e, d);
                dirs.push_back(e);
            }
            m["Total"] = nd;
            m["Directions"] = dirs;
        }
        return m;
    }
    void PrintJsonReports(const std::string& stem)
    {
        {
            std::ofstream os(stem + ".adj.jsonl");
            JsonHeader(os, "adj");
            Json st;
            st["iteration"] = report_mode_ ? last_iterations_ : (uint32_t)iterations_.size();
            st["unknown_parameters"] = stats_.unknown_params;
            st["measurement_params"] = stats_.measurement_params;
            st["potential_outliers"] = stats_.outliers;
            st["dof"] = (long long)stats_.dof;
            st["chisq"] = stats_.chi_squared;
            st["sigma_zero"] = stats_.sigma_zero;
            st["global_pelzer"] = stats_.global_pelzer;
            st["chisq_lower"] = chiLower_;
            st["chisq_upper"] = chiUpper_;
            st["confidence_interval"] = a_.confidence_interval;
            st["chisq_test"] = stats_.dof < 1 ? "no_redundancy" : (passFail_ == 0 ? "passed" : (passFail_ == 1 ? "warning" : "failed"));
            WriteRecord(os, "DnaStatistics", st);
            if (a_.output_adj_msr) {
                std::vector<uint32_t> list = CollectMeasurements(nullptr, -1, false);
                SortMeasurements(list);
                for (uint32_t f : list)
                    WriteRecord(os, "DnaMeasurement", JsonMeasurement(f));
            }
            for (uint32_t i : StationOrder(nullptr))
                WriteRecord(os, "DnaStation", JsonAdjustedStation(i));
        }
        {
            std::ofstream os(stem + ".xyz.jsonl");
            JsonHeader(os, "xyz");
            for (uint32_t i : StationOrder(nullptr))
                WriteRecord(os, "DnaStation", JsonAdjustedStation(i));
        }
        if (a_.output_pos_uncertainty) {
            std::ofstream os(stem + ".apu.jsonl");
            JsonHeader(os, "apu");
            for (uint32_t i : StationOrder(nullptr)) {
                Json s = JsonStationIdentity(stn_[i]);
                s["Uncertainty"] = JsonUncertainty(i, true);
                WriteRecord(os, "DnaStation", s);
            }
        }
        if (a_.output_corrections) {
            std::ofstream os(stem + ".cor.jsonl");
            JsonHeader(os, "cor");
            for (size_t i = 0; i < stn_.size(); ++i) {
                const dna_stn_t& s = stn_[i];
                double o[3], R[9];
                OriginalXYZ(i, o);
                local_rotation(s.currentLatitude, s.currentLongitude, R);
                const double d[3] = {est_[3 * i] - o[0], est_[3 * i + 1] - o[1], est_[3 * i + 2] - o[2]};
                Json j = JsonStationIdentity(s), c;
                j["Initial"] = JsonInitial(s);
                c["dE"] = R[0] * d[0] + R[3] * d[1] + R[6] * d[2];
                c["dN"] = R[1] * d[0] + R[4] * d[1] + R[7] * d[2];
                c["dUp"] = R[2] * d[0] + R[5] * d[1] + R[8] * d[2];
                j["Corrections"] = c;
                WriteRecord(os, "DnaStation", j);
            }
        }
    }

    // ---- iteration diagnostics (UpdateIterationDiagnostics ADJ:7450-7547, PrintOscillationSummary ADJ:7549-7610,
    // PrintSuspectMeasurementSummary ADJ:7652-7779): a station whose correction vector flips direction with a similar
    // magnitude on successive iterations (cosine < -0.5, ratio 0.3-3) for two iterations running is oscillating; the
    // summary names the worst, and the measurements that touch them or exceed the critical n-statistic
    struct OscillationRecord {
        uint32_t stn, firstIteration, lastIteration, maxCycles;
        double firstMag, lastMag, lastE, lastN, lastUp;
    };
    void UpdateIterationDiagnostics()
    {
        std::vector<double> corr(3 * stn_.size());
        check(gadj_get_corrections(ctx_, corr.data()));
        const uint32_t it = (uint32_t)iterations_.size();
        if (corrPrev_.empty()) {
            corrPrev_ = corr;
            stnOscCount_.assign(stn_.size(), 0);
            return;
        }
        for (size_t s = 0; s < stn_.size(); ++s) {
            const double* c = &corr[3 * s];
            const double* p = &corrPrev_[3 * s];
            const double magCurr = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]), magPrev = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
            if (magCurr < 0.001 && magPrev < 0.001) {   // sub-millimetre
                stnOscCount_[s] = 0;
                continue;
            }
            const double denom = magCurr * magPrev;
            const double cosAngle = denom > 1e-30 ? (c[0] * p[0] + c[1] * p[1] + c[2] * p[2]) / denom : 0.0;
            const double ratio = magPrev > 1e-30 ? magCurr / magPrev : 0.0;
            if (cosAngle < -0.5 && ratio > 0.3 && ratio < 3.0)
                stnOscCount_[s]++;
            else
                stnOscCount_[s] = 0;
            if (stnOscCount_[s] < 2)
                continue;
            double R[9];
            local_rotation(stn_[s].currentLatitude, stn_[s].currentLongitude, R);
            const double e = R[0] * c[0] + R[3] * c[1] + R[6] * c[2], n = R[1] * c[0] + R[4] * c[1] + R[7] * c[2],
                         u = R[2] * c[0] + R[5] * c[1] + R[8] * c[2];
            const double mag = std::sqrt(e * e + n * n + u * u);
            auto hit = oscHistory_.find((uint32_t)s);
            if (hit == oscHistory_.end())
                oscHistory_[(uint32_t)s] = OscillationRecord{(uint32_t)s, it, it, stnOscCount_[s], mag, mag, e, n, u};
            else {
                hit->second.lastIteration = it;
                hit->second.maxCycles = stnOscCount_[s];
                hit->second.lastMag = mag;
                hit->second.lastE = e, hit->second.lastN = n, hit->second.lastUp = u;
            }
        }
        corrPrev_ = corr;
    }
    void PrintOscillationSummary(std::ostream& os) const
    {
        std::vector<const OscillationRecord*> sorted;
        for (const auto& kv : oscHistory_)
            if (std::max(kv.second.firstMag, kv.second.lastMag) >= 0.1)
                sorted.push_back(&kv.second);
        if (sorted.empty())
            return;
        std::sort(sorted.begin(), sorted.end(), [](const OscillationRecord* a, const OscillationRecord* b) {
            return std::max(a->firstMag, a->lastMag) > std::max(b->firstMag, b->lastMag);
        });
        const size_t limit = std::min<size_t>(sorted.size(), 20);
        os << "\n+ Oscillating stations detected (" << sorted.size() << " total, showing top " << limit << "):\n";
        for (size_t i = 0; i < limit; ++i) {
            const OscillationRecord* r = sorted[i];
            const double hz = std::hypot(r->lastE, r->lastN), vt = std::fabs(r->lastUp);
            const char* dir = vt < 0.01 * hz ? "horizontal" : (hz < 0.01 * vt ? "vertical" : "3D");
            os << "  - " << stn_[r->stn].stationName << std::fixed << std::setprecision(1) << " - " << r->firstMag << "m to " << r->lastMag << "m, " << dir
               << ", " << r->maxCycles << " cycles (iterations " << r->firstIteration << "-" << r->lastIteration << ")\n";
        }
        os.unsetf(std::ios::floatfield);
    }
    void PrintSuspectMeasurementSummary(std::ostream& os, size_t limit = 20) const
    {
        struct Suspect {
            uint32_t rec;
            double absN;
            bool critical, osc;
        };
        std::vector<Suspect> oscList, outList;
        const double crit = stats_.critical_value;
        for (uint32_t i = 0; i < msr_.size(); ++i) {
            const dna_msr_t& m = msr_[i];
            if (m.ignore || !std::isfinite(m.NStat) || !std::isfinite(m.residualPrec) || m.residualPrec <= 0.0)
                continue;
            if ((m.measType == 'G' || m.measType == 'X' || m.measType == 'Y') && m.measStart > 2)
                continue;   // covariance records carry no statistics
            const bool critical = std::fabs(m.NStat) > crit;
            bool osc = oscHistory_.count(m.station1) > 0;
            if (!osc && m.measurementStations >= 2 && m.measType != 'Y')
                osc = oscHistory_.count(m.station2) > 0;
            if (!osc && m.measurementStations >= 3 && m.measType == 'A')
                osc = oscHistory_.count(m.station3) > 0;
            if (osc)
                oscList.push_back({i, std::fabs(m.NStat), critical, true});
            else if (critical)
                outList.push_back({i, std::fabs(m.NStat), true, false});
        }
        auto by_n = [](const Suspect& a, const Suspect& b) { return a.absN == b.absN ? a.rec < b.rec : a.absN > b.absN; };
        std::sort(oscList.begin(), oscList.end(), by_n);
        std::sort(outList.begin(), outList.end(), by_n);
        auto print = [&](const char* title, const std::vector<Suspect>& list) {
            if (list.empty())
                return;
            const size_t n = std::min(list.size(), limit);
            os << "\n+ " << title << " (" << list.size() << " total, showing top " << n << "):\n";
            char buf[512];
            for (size_t k = 0; k < n; ++k) {
                const dna_msr_t& m = msr_[list[k].rec];
                std::string names = stn_[m.station1].stationName;
                if (m.measurementStations >= 2 && m.measType != 'Y')
                    names += std::string(" -> ") + stn_[m.station2].stationName;
                if (m.measurementStations >= 3 && m.measType == 'A')
                    names += std::string(" -> ") + stn_[m.station3].stationName;
                snprintf(buf, sizeof(buf), "  - %c msr %u cluster %u file-order %u %s: N=%.2f", m.measType, list[k].rec, m.clusterID, m.fileOrder, names.c_str(),
                         m.NStat);
                os << buf;
                if (std::isfinite(m.TStat) && std::fabs(m.TStat) > 0.0) {
                    snprintf(buf, sizeof(buf), ", T=%.2f", m.TStat);
                    os << buf;
                }
                snprintf(buf, sizeof(buf), ", corr=%.3e, residual precision=%.3e, Pelzer=%.2f", m.measCorr, m.residualPrec, m.PelzerRel);
                os << buf << (list[k].critical ? ", exceeds critical" : "") << (list[k].osc ? ", touches oscillating station" : "") << "\n";
            }
        };
        print("Suspect measurements connected to oscillating stations", oscList);
        print(oscList.empty() ? "Largest measurement N-statistics" : "Largest remaining measurement N-statistics", outList);
    }

    // An adjustment that ran out of iterations reports its iterations and status only (WRAP:1386-1390): no statistics
    void PrintFailedAdjustment()
    {
        const std::string stem = a_.output_folder + "/" + a_.network_name + "." + ModeSuffix();
        std::ofstream adj(stem + ".adj");
        PrintOutputFileHeaderInfo(adj, "DYNADJUST ADJUSTMENT OUTPUT FILE", stem + ".adj");
        adj << "\n+ Initialising adjustment\n+ Loading network files\n+ Allocating memory\n\n+ Preparing for adjustment...  done.\n";
        adj << "+ Commencing " << (a_.adjust_mode == SimultaneousMode ? "simultaneous" : "phased") << " adjustment\n\n";
        for (size_t i = 0; i < iterations_.size(); ++i) {
            PrintIteration(adj, (uint32_t)i + 1, iterations_[i]);
            adj << iter_pre_[i] << iter_post_[i];
        }
        const std::string dash(80, '-');
        adj << "\n" << dash << "\n" << std::left << std::setw(35) << "SOLUTION" << "Failed to converge\n";
        char buf[64];
        snprintf(buf, sizeof(buf), "00:00:%09.6f", total_ms_ / 1e3);
        adj << std::left << std::setw(35) << "Total time" << buf << "\n\n";
        std::ofstream xyz(stem + ".xyz");
        PrintOutputFileHeaderInfo(xyz, "DYNADJUST COORDINATE OUTPUT FILE", stem + ".xyz");
    }

    // ---- DNA / DynaML exports of the adjusted stations (PrintEstimatedStationCoordinatestoDNAXML PRN:2775-2903;
    // WriteDNAStn / WriteDynaMLStn dnastation.cpp:825-886): <adj file>.stn and <adj file>.stn.xml, stations in the order
    // of the imported file, coordinates in the form they were supplied in (LLH / UTM: orthometric height)
    static std::string today_ddmmyyyy()
    {
        std::time_t t = std::time(nullptr);
        std::tm tmv;
        localtime_r(&t, &tmv);
        char b[32];
        std::strftime(b, sizeof(b), "%d.%m.%Y", &tmv);
        return b;
    }
    void dna_header(std::ostream& os, const char* type, size_t count) const
    {   // dnastringfuncs.cpp:230-258
        os << "!#=DNA 3.01 " << type << std::setw(14) << std::right << today_ddmmyyyy() << std::setw(14) << frame_name() << std::setw(14)
           << bst_meta_.epoch << std::setw(10) << count << "\n"
           << "* Created by:   dnaadjust (dynadjust_b200), B200 geodetic adjustment. \n* Version:      1.0. \n";
    }
    void dynaml_header(std::ostream& os, const char* type) const
    {   // dnastringfuncs.cpp:173-190
        os << "<?xml version=\"1.0\"?>\n<DnaXmlFormat type=\"" << type << "\" referenceframe=\"" << frame_name() << "\" epoch=\"" << bst_meta_.epoch
           << "\" xmlns:xsi=\"http://www.w3.org/2001/XMLSchema-instance\" xsi:noNamespaceSchemaLocation=\"DynaML.xsd\">\n"
           << "<!-- Created by:   dnaadjust (dynadjust_b200), B200 geodetic adjustment -->\n<!-- Version:      1.0 -->\n";
    }
    static std::string xml_escape(const char* s)
    {
        std::string o;
        for (; *s; ++s)
            o += *s == '&' ? "&amp;" : *s == '<' ? "&lt;" : *s == '>' ? "&gt;" : std::string(1, *s);
        return o;
    }

    void PrintEstimatedStationCoordinatestoDNAXML(const std::string& file, bool dynaml, const std::string& adj_file) const
    {
        std::ofstream os(file);
        const std::string source = "Source data:  Coordinates estimated from least squares adjustment.";
        if (dynaml) {
            dynaml_header(os, "Station File");
            os << "<!-- File type:    Station file -->\n<!-- Project name: " << a_.network_name << " -->\n<!-- " << source << " -->\n<!-- Adj file:     "
               << adj_file << " -->\n";
        } else {
            dna_header(os, "STN", stn_.size());
            os << "* File type:    Station file\n* Project name: " << a_.network_name << "\n* " << source << "\n* Adj file:     " << adj_file << "\n";
        }
        std::vector<uint32_t> list;
        if (a_.adjust_mode == Phased_Block_1Mode && !seg_.isl.empty()) {
            list = seg_.isl[0];
            if (!seg_.jsl.empty())
                list.insert(list.end(), seg_.jsl[0].begin(), seg_.jsl[0].end());
        } else {
            list.resize(stn_.size());
            for (size_t i = 0; i < list.size(); ++i)
                list[i] = (uint32_t)i;
        }
        std::stable_sort(list.begin(), list.end(), [&](uint32_t a, uint32_t b) { return stn_[a].fileOrder < stn_[b].fileOrder; });
        const gadj::Ellipsoid ell = Ellipsoid();
        for (uint32_t i : list) {
            const dna_stn_t& s = stn_[i];
            const char* type = "LLH";
            double c[3] = {s.currentLatitude, s.currentLongitude, s.currentHeight};
            std::string zone;
            int p12 = 4;
            switch (s.suppliedStationType) {
            case DNA_XYZ_TYPE:
                type = "XYZ";
                gadj::geo_to_cart(ell, s.currentLatitude, s.currentLongitude, s.currentHeight, c);
                break;
            case DNA_UTM_TYPE: {
                type = "UTM";
                double z;
                GeoToGrid(ell, s.currentLatitude, s.currentLongitude, &c[0], &c[1], &z);
                c[2] -= s.geoidSep;
                zone = std::to_string((int)z);
                break;
            }
            case DNA_LLh_TYPE:
                type = "LLh";
                [[fallthrough]];
            default:   // LLH (and ENU, which the reference writes as LLH)
                if (s.suppliedStationType != DNA_LLh_TYPE)
                    c[2] -= s.geoidSep;
                c[0] = std::atof(hp_dms(s.currentLatitude, 14).c_str());
                c[1] = std::atof(hp_dms(s.currentLongitude, 14).c_str());
                p12 = 10;
            }
            char cst[4] = {s.stationConst[0], s.stationConst[1], s.stationConst[2], 0};
            if (dynaml) {
                os << "  <DnaStation>\n    <Name>" << xml_escape(s.stationName) << "</Name>\n    <Constraints>" << cst << "</Constraints>\n    <Type>" << type
                   << "</Type>\n    <StationCoord>\n      <Name>" << xml_escape(s.stationName) << "</Name>\n      <XAxis>" << Fixed(c[0], 0, p12)
                   << "</XAxis>\n      <YAxis>" << Fixed(c[1], 0, p12) << "</YAxis>\n      <Height>" << Fixed(c[2], 0, 4) << "</Height>\n";
                if (!zone.empty())
                    os << "      <HemisphereZone>" << zone << "</HemisphereZone>\n";
                os << "    </StationCoord>\n    <Description>" << xml_escape(s.description) << "</Description>\n  </DnaStation>\n";
            } else {
                os << std::left << std::setw(20) << s.stationName << std::setw(3) << cst << " " << std::setw(3) << type << std::right << Fixed(c[0], 20, p12)
                   << Fixed(c[1], 20, p12) << Fixed(c[2], 20, 4) << std::setw(3) << (zone.empty() ? " " : zone) << " " << s.description << "\n";
            }
        }
        if (dynaml)
            os << "</DnaXmlFormat>\n";
    }

    // ---- DNA / DynaML exports of the estimates as GNSS point clusters (PrintEstimatedStationCoordinatestoDNAXML_Y
    // PRN:3012-3164; CDnaGpsPoint::WriteDNAMsr / WriteDynaMLMsr dnagpspoint.cpp:232-366): one Y cluster per block — the
    // Cartesian estimates of its stations with the block's full variance matrix — in <adj file>.msr / .msr.xml
    void PrintEstimatedStationCoordinatestoDNAXML_Y(const std::string& file, bool dynaml, const std::string& adj_file)
    {
        std::ofstream os(file);
        const uint32_t nblocks = (uint32_t)info_.nfronts;
        std::ostringstream src;
        src << "Source data:  Coordinates and uncertainties for " << stn_.size() << " unique stations in " << nblocks
            << " blocks estimated from least squares adjustment.";
        if (dynaml) {
            dynaml_header(os, "Measurement File");
            os << "<!-- File type:    Measurement file -->\n<!-- Project name: " << a_.network_name << " -->\n<!-- " << src.str()
               << " -->\n<!-- Adj file:     " << adj_file << " -->\n";
        } else {
            dna_header(os, "MSR", nblocks);
            os << "* File type:    Measurement file\n* Project name: " << a_.network_name << "\n* " << src.str() << "\n* Adj file:     " << adj_file << "\n";
        }
        const std::string frame = frame_name(), epoch = bst_meta_.epoch;
        char num[64];
        auto sci = [&](double v) {
            snprintf(num, sizeof(num), dynaml ? "%.13e" : "%20.13e", v);
            return std::string(num);
        };
        for (uint32_t b = 0; b < nblocks; ++b) {
            if (a_.adjust_mode == Phased_Block_1Mode && b > 0)
                break;
            uint32_t n = 0;
            check(gadj_get_block_vcv(ctx_, b, &n, nullptr, 0, nullptr));
            std::vector<uint32_t> st(n);
            const size_t dim = 3 * (size_t)n;
            std::vector<double> q(dim * (dim + 1) / 2);
            check(gadj_get_block_vcv(ctx_, b, &n, st.data(), n, q.data()));
            auto at = [&](size_t i, size_t j) { return i >= j ? q[j * dim - j * (j - 1) / 2 + (i - j)] : q[i * dim - i * (i - 1) / 2 + (j - i)]; };
            if (dynaml) {
                os << "  <!--\n    - Estimated station coordinates and uncertainties";
                if (nblocks > 1)
                    os << " for block " << b + 1;
                os << "\n    - Type (Y) GPS point cluster (set of " << n << " stations)\n  -->\n";
                os << "  <DnaMeasurement>\n    <Type>Y</Type>\n    <Source></Source>\n    <Ignore/>\n    <ReferenceFrame>" << frame << "</ReferenceFrame>\n    <Epoch>" << epoch
                   << "</Epoch>\n    <Vscale>1.000</Vscale>\n    <Pscale>1.000</Pscale>\n    <Lscale>1.000</Lscale>\n    <Hscale>1.000</Hscale>\n    <Coords>XYZ</Coords>\n"
                   << "    <Total>" << n << "</Total>\n";
            }
            for (uint32_t k = 0; k < n; ++k) {
                const double* x = &est_[3 * (size_t)st[k]];
                const size_t r = 3 * (size_t)k;
                if (dynaml) {
                    os << "    <First>" << xml_escape(stn_[st[k]].stationName) << "</First>\n    <Clusterpoint>\n      <X>" << Fixed(x[0], 0, 4) << "</X>\n      <Y>"
                       << Fixed(x[1], 0, 4) << "</Y>\n      <Z>" << Fixed(x[2], 0, 4) << "</Z>\n      <SigmaXX>" << sci(at(r, r)) << "</SigmaXX>\n      <SigmaXY>"
                       << sci(at(r, r + 1)) << "</SigmaXY>\n      <SigmaXZ>" << sci(at(r, r + 2)) << "</SigmaXZ>\n      <SigmaYY>" << sci(at(r + 1, r + 1))
                       << "</SigmaYY>\n      <SigmaYZ>" << sci(at(r + 1, r + 2)) << "</SigmaYZ>\n      <SigmaZZ>" << sci(at(r + 2, r + 2)) << "</SigmaZZ>\n";
                    for (uint32_t j = k + 1; j < n; ++j) {
                        os << "      <PointCovariance>\n";
                        static const char* tag[9] = {"m11", "m12", "m13", "m21", "m22", "m23", "m31", "m32", "m33"};
                        for (int a = 0; a < 3; ++a)
                            for (int c = 0; c < 3; ++c)
                                os << "        <" << tag[3 * a + c] << ">" << sci(at(r + a, 3 * (size_t)j + c)) << "</" << tag[3 * a + c] << ">\n";
                        os << "      </PointCovariance>\n";
                    }
                    os << "    </Clusterpoint>\n";
                    continue;
                }
                os << "Y " << std::left << std::setw(20) << stn_[st[k]].stationName;
                if (k == 0)
                    os << std::setw(20) << "XYZ" << std::setw(20) << n << std::right << Fixed(1.0, 10, 2) << Fixed(1.0, 10, 2) << Fixed(1.0, 10, 2)
                       << Fixed(1.0, 10, 2) << std::setw(20) << frame << std::setw(20) << epoch;
                os << "\n" << std::string(62, ' ') << Fixed(x[0], 20, 4) << sci(at(r, r)) << "\n"
                   << std::string(62, ' ') << Fixed(x[1], 20, 4) << sci(at(r, r + 1)) << sci(at(r + 1, r + 1)) << "\n"
                   << std::string(62, ' ') << Fixed(x[2], 20, 4) << sci(at(r, r + 2)) << sci(at(r + 1, r + 2)) << sci(at(r + 2, r + 2)) << "\n";
                for (uint32_t j = k + 1; j < n; ++j)
                    for (int a = 0; a < 3; ++a)
                        os << std::string(82, ' ') << sci(at(r + a, 3 * (size_t)j)) << sci(at(r + a, 3 * (size_t)j + 1)) << sci(at(r + a, 3 * (size_t)j + 2)) << "\n";
            }
            if (dynaml)
                os << "  </DnaMeasurement>\n";
        }
        if (dynaml)
            os << "</DnaXmlFormat>\n";
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
    void PrintPositionalUncertainty(const std::string& file)
    {
        std::ofstream os(file);
        PrintStationFileHeader(os, "POSITIONAL UNCERTAINTY", file);
        auto var = [&](const char* n, const std::string& v) { os << std::left << std::setw(35) << n << v << "\n"; };
        var("PU confidence interval:", "95.0%");
        var("Error ellipse axes:", "68.3% (1 sigma)");
        var("Variances:", "68.3% (1 sigma)");
        var("Stations printed in blocks:", "No");
        var("Variance matrix units:", a_.apu_vcv_enu ? "ENU" : "XYZ");
        var("Full covariance matrix:", a_.output_pu_covariances ? "Yes" : "No");
        if (!a_.type_b_global.empty())
            var("Type B uncertainties:", a_.type_b_global);
        if (!a_.type_b_file.empty())
            var("Type B uncertainty file:", a_.type_b_file);
        os << std::string(80, '-') << "\n\n";
        os << "Positional uncertainty of adjusted station coordinates\n";
        os << "------------------------------------------------------\n\n";
        const char* vn = a_.apu_vcv_enu ? "enu" : "XYZ";
        char v1[16], v2[16], v3[16];
        snprintf(v1, sizeof(v1), "Variance(%c)", vn[0]);
        snprintf(v2, sizeof(v2), "Variance(%c)", vn[1]);
        if (a_.apu_vcv_enu)
            snprintf(v3, sizeof(v3), "Variance(up)");
        else
            snprintf(v3, sizeof(v3), "Variance(Z)");
        char head[512];
        snprintf(head, sizeof(head), "%-20s%2s%14s%15s%11s%11s%13s%13s%13s%19s%19s%19s", "Station", "", "Latitude", "Longitude", "Hz PosU",
                 "Vt PosU", "Semi-major", "Semi-minor", "Orientation", v1, v2, v3);
        const std::string header = std::string(head) + "\n" + std::string(20 + 2 + 14 + 15 + 11 + 11 + 13 + 13 + 13 + 19 + 19 + 19, '-') + "\n";
        if (!a_.output_pu_covariances) {
            os << header;
            for (uint32_t i : StationOrder(nullptr))
                PrintPosUncertainty(os, i);
            return;
        }
        // --output-all-covariances (PrintPosUncertainty PRN:4438-4484): after each station, its 3x3 covariance blocks with the
        // stations that follow it in the block, from the block's dense variance matrix; phased adjustments list block by block
        const uint32_t nblocks = (uint32_t)info_.nfronts;
        for (uint32_t b = 0; b < nblocks; ++b) {
            if (a_.adjust_mode == Phased_Block_1Mode && b > 0)
                break;
            uint32_t n = 0;
            check(gadj_get_block_vcv(ctx_, b, &n, nullptr, 0, nullptr));
            std::vector<uint32_t> st(n);
            const size_t dim = 3 * (size_t)n;
            std::vector<double> q(dim * (dim + 1) / 2);
            check(gadj_get_block_vcv(ctx_, b, &n, st.data(), n, q.data()));
            auto at = [&](size_t i, size_t j) { return i >= j ? q[j * dim - j * (j - 1) / 2 + (i - j)] : q[i * dim - i * (i - 1) / 2 + (j - i)]; };
            std::vector<uint32_t> order(n);   // positions in the block, in the order the stations are listed
            for (uint32_t k = 0; k < n; ++k)
                order[k] = k;
            std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
                return a_.sort_stn_orig_order ? stn_[st[x]].fileOrder < stn_[st[y]].fileOrder : st[x] < st[y];
            });
            if (a_.adjust_mode != SimultaneousMode)
                os << "Block " << b + 1 << "\n";
            os << header;
            const int pad = 2 + 14 + 15 + 11 + 11 + 13 + 13 + 13;
            char buf[256];
            for (uint32_t k = 0; k < n; ++k) {
                const uint32_t i = st[order[k]];
                PrintPosUncertainty(os, i);
                double R[9];
                local_rotation(stn_[i].currentLatitude, stn_[i].currentLongitude, R);
                for (uint32_t m = k + 1; m < n; ++m) {
                    double c[9], cl[9];
                    for (int x = 0; x < 3; ++x)
                        for (int y = 0; y < 3; ++y)
                            c[3 * x + y] = at(3 * (size_t)order[k] + x, 3 * (size_t)order[m] + y);
                    const double* v = c;
                    if (a_.apu_vcv_enu) {
                        rotate_sym(R, c, cl);
                        v = cl;
                    }
                    for (int x = 0; x < 3; ++x) {
                        snprintf(buf, sizeof(buf), "%-20s%*s%19.9e%19.9e%19.9e", x == 0 ? stn_[st[order[m]]].stationName : "", pad, "", v[3 * x],
                                 v[3 * x + 1], v[3 * x + 2]);
                        os << buf << "\n";
                    }
                }
            }
            os << "\n";
        }
    }

    void PrintPosUncertainty(std::ostream& os, size_t i) const
    {
        char buf[512];
        const int pad = 20 + 2 + 14 + 15 + 11 + 11 + 13 + 13 + 13;
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
            double o[3];
            OriginalXYZ(i, o);
            const double d[3] = {est_[3 * i] - o[0], est_[3 * i + 1] - o[1], est_[3 * i + 2] - o[2]};
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
                     AngleString(az, 0, 0, 0).c_str(), AngleString(va, 0, 0, 0).c_str(), sd, hd, e, n, u);   // "ddd mm ss", carries exact
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
        // discontinuity sites (AddDiscontinuitySites LDR:314-359): when dnaimport renamed stations of a discontinuity file
        // (stationName differs from stationNameOrig), a constraint given for the original name also goes to its renamed
        // sites, and a name that no longer exists is passed over instead of being an error (LDR:246-249)
        bool discontinuities = false;
        for (const dna_stn_t& st : stn_)
            if (st.stationNameOrig[0] && std::strncmp(st.stationName, st.stationNameOrig, sizeof(st.stationName)) != 0)
                discontinuities = true;
        if (discontinuities) {
            const size_t given = tok.size();
            for (size_t k = 0; k + 1 < given; k += 2)
                for (const dna_stn_t& st : stn_)
                    if (tok[k] == st.stationNameOrig && tok[k] != st.stationName) {
                        tok.push_back(st.stationName);
                        tok.push_back(tok[k + 1]);
                    }
        }
        for (size_t k = 0; k + 1 < tok.size(); k += 2) {
            std::string c = tok[k + 1];
            for (char& ch : c)
                ch = (char)std::toupper((unsigned char)ch);
            auto it = by_name.find(tok[k]);
            if (it == by_name.end()) {
                if (discontinuities)
                    continue;
                SignalExceptionAdjustment("The supplied constraint station '" + tok[k] + "' is not in the stations map");
            }
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

    // degrees.minutes-seconds "HP" notation ddd.mmsss.. with `decimals` places (RadtoDms + fixed, PRN:1632-1650); integer
    // arithmetic on the last printed place of a second so that carries are exact
    static std::string hp_dms(double rad, int decimals = 9)
    {
        const int sp = std::max(0, decimals - 4);
        long long scale = 1;
        for (int k = 0; k < sp; ++k)
            scale *= 10;
        const double deg = std::fabs(rad) * 180.0 / 3.14159265358979323846;
        const long long units = std::llround(deg * 3600.0 * (double)scale);
        const long long d = units / (3600LL * scale), rem = units % (3600LL * scale);
        const long long mi = rem / (60LL * scale), sec = rem % (60LL * scale);
        char b[64], frac[32] = "";
        if (sp > 0)
            snprintf(frac, sizeof(frac), "%0*lld", sp, sec % scale);
        snprintf(b, sizeof(b), "%s%lld.%02lld%02lld%s", rad < 0 && units > 0 ? "-" : "", d, mi, sec / scale, frac);
        return b;
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
        var("Reference frame:", frame_name());
        var("Epoch:", bst_meta_.epoch);
        var("Geoid model:", "");
        if (a_.adjust_mode != SimultaneousMode)
            var("Segmentation file:", a_.seg_file.empty() ? a_.input_folder + "/" + a_.network_name + ".seg" : a_.seg_file);
        std::ostringstream t;
        t << a_.fixed_std_dev;
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
        if (!a_.station_constraints.empty())
            var("Station constraints:", a_.station_constraints);   // PRN:3490
        var("Station coordinate types:", a_.stn_coord_types);
        var("Stations printed in blocks:", a_.adjust_mode != SimultaneousMode && a_.output_stn_blocks ? "Yes" : "No");
        if (a_.stn_corrections)
            var("Station coordinate corrections:", "Yes");
        if (!a_.type_b_global.empty())
            var("Type B uncertainties:", a_.type_b_global);
        if (!a_.type_b_file.empty())
            var("Type B uncertainty file:", a_.type_b_file);
        // user comments, wrapped at word breaks to the value column ("\n" in the text starts a new line) (PRN:3540-3592)
        if (!a_.comments.empty()) {
            std::string text = a_.comments, label = "Comments: ";
            for (size_t p2; (p2 = text.find("\\n")) != std::string::npos;)
                text.replace(p2, 2, "\n");
            std::istringstream lines(text);
            for (std::string ln; std::getline(lines, ln);) {
                while (!ln.empty() && ln[0] == ' ')
                    ln.erase(0, 1);
                while (ln.size() > 45) {
                    size_t cut = ln.rfind(' ', 45);
                    if (cut == std::string::npos || cut == 0)
                        cut = 45;
                    var(label.c_str(), ln.substr(0, cut));
                    label = " ";
                    ln.erase(0, cut);
                    while (!ln.empty() && ln[0] == ' ')
                        ln.erase(0, 1);
                }
                var(label.c_str(), ln);
                label = " ";
            }
        }
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
        var("SOLUTION") << (report_mode_ ? "Printing results of last adjustment only"
                                         : (adjustStatus_ == ADJUST_SUCCESS ? "Converged" : "Failed to converge after maximum iterations")) << "\n";
        char buf[64];
        snprintf(buf, sizeof(buf), "00:00:%09.6f", total_ms_ / 1e3);
        var("Total time") << buf << "\n\n";
        PrintStatisticsSummary(os, true);
    }

    void PrintStatisticsSummary(std::ostream& os, bool printPelzer) const
    {
        auto var = [&](const char* n) -> std::ostream& { return os << std::left << std::setw(35) << n; };
        var("Number of unknown parameters") << stats_.unknown_params << "\n";
        var("Number of measurements") << stats_.measurement_params;
        if (stats_.outliers > 0)
            os << "  (" << stats_.outliers << " potential outlier" << (stats_.outliers > 1 ? "s" : "") << ")";
        os << "\n";
        var("Degrees of freedom") << stats_.dof << "\n";
        var("Chi squared") << std::fixed << std::setprecision(2) << stats_.chi_squared << "\n";
        var("Rigorous Sigma Zero") << std::fixed << std::setprecision(3) << stats_.sigma_zero << "\n";
        if (printPelzer)
            var("Global (Pelzer) Reliability") << std::fixed << std::setprecision(3) << stats_.global_pelzer
                                               << "   (excludes non redundant measurements)\n";
        os << "\n";
        if (a_.adjust_mode == Phased_Block_1Mode) {   // no global test in block-1 mode (ADJ:7140-7147)
            os << "\n";
            return;
        }
        std::ostringstream t;
        t << "Chi-Square test (" << std::fixed << std::setprecision(1) << a_.confidence_interval << "%)";
        var(t.str().c_str()) << std::fixed << std::setprecision(3) << chiLower_ << " < " << stats_.sigma_zero << " < " << chiUpper_
                             << "          " << (stats_.dof < 1 ? "NO REDUNDANCY" : (passFail_ == 0 ? "*** PASSED ***" : (passFail_ == 1 ? "*** WARNING ***" : "*** FAILED ***")))
                             << "\n\n";
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

    // ---- adjusted measurements table (PrintAdjMeasurements PRN:1682-1782, PrintMeasurementRecords PRN:2025-2117) ------
    // Measurements are listed by their first record (a G baseline, an X / Y cluster, a direction set, a scalar row),
    // sorted as --sort-adj-msr-field asks, and printed by type.
    size_t MeasurementSpan(size_t i) const
    {
        const dna_msr_t& m = msr_[i];
        switch (m.measType) {
        case 'G': case 'X': case 'Y': {
            size_t j = i;
            const uint32_t count = std::max<uint32_t>(1u, m.vectorCount1);
            for (uint32_t k = 0; k < count && j < msr_.size(); ++k)
                j += 3 + 3 * (size_t)msr_[j].vectorCount2;
            return std::min(j, msr_.size()) - i;
        }
        case 'D':
            return std::max<uint32_t>(1u, m.vectorCount1);
        default:
            return 1;
        }
    }

    std::vector<uint32_t> CollectMeasurements(const std::vector<int32_t>* rec_block, int32_t block, bool ignored) const
    {
        std::vector<uint32_t> list;
        for (size_t i = 0; i < msr_.size();) {
            const size_t span = MeasurementSpan(i);
            if ((msr_[i].ignore != 0) == ignored && (!rec_block || (*rec_block)[i] == block))
                list.push_back((uint32_t)i);
            i += span;
        }
        return list;
    }

    // largest |field| over the components of a compound measurement (CompareMeas*_PairFirst, dnatemplatestnmsrfuncs.hpp:1148-1810)
    template <typename F>
    double LargestOf(uint32_t first, F field) const
    {
        const dna_msr_t& m = msr_[first];
        double v = 0.0;
        switch (m.measType) {
        case 'G': case 'X': case 'Y': {
            size_t j = first;
            const uint32_t count = std::max<uint32_t>(1u, m.vectorCount1);
            for (uint32_t k = 0; k < count && j + 2 < msr_.size(); ++k) {
                for (int q = 0; q < 3; ++q)
                    v = std::max(v, std::fabs(field(msr_[j + q])));
                j += 3 + 3 * (size_t)msr_[j].vectorCount2;
            }
            return v;
        }
        case 'D':
            for (uint32_t d = 0; d < std::max<uint32_t>(1u, m.vectorCount1) && first + d < msr_.size(); ++d)
                v = std::max(v, std::fabs(field(msr_[first + d])));
            return v;
        default:
            return std::fabs(field(m));
        }
    }

    void SortMeasurements(std::vector<uint32_t>& list) const
    {
        auto by_keys = [&](auto key) {
            std::stable_sort(list.begin(), list.end(), [&](uint32_t a, uint32_t b) { return key(msr_[a]) < key(msr_[b]); });
        };
        auto by_largest = [&](auto field) {
            std::vector<std::pair<double, uint32_t>> k;
            for (uint32_t f : list)
                k.emplace_back(LargestOf(f, field), f);
            std::stable_sort(k.begin(), k.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
            for (size_t i = 0; i < k.size(); ++i)
                list[i] = k[i].second;
        };
        switch (a_.sort_adj_msr) {
        case 1:   // measurement type, first station, second station, value
            by_keys([](const dna_msr_t& m) { return std::make_tuple(m.measType, m.station1, m.station2, m.term1); });
            break;
        case 2:   // "instrument station" sorts on the second station (SortMeasurementsbyToStn, PRN:1721)
            by_keys([](const dna_msr_t& m) { return std::make_tuple(m.station2, m.measType, m.station1, m.term1); });
            break;
        case 3:   // "target station" sorts on the first station (SortMeasurementsbyFromStn, PRN:1724)
            by_keys([](const dna_msr_t& m) { return std::make_tuple(m.station1, m.measType, m.station2, m.term1); });
            break;
        case 4: by_largest([](const dna_msr_t& m) { return m.term1; }); break;
        case 5: by_largest([](const dna_msr_t& m) { return m.measCorr; }); break;
        case 6: by_largest([](const dna_msr_t& m) { return m.measAdjPrec; }); break;
        case 7: by_largest([](const dna_msr_t& m) { return m.NStat; }); break;
        default: break;   // original (file) order
        }
    }

    // StringFromTW (dnastrmanipfuncs.hpp:216-264): fixed notation when it fits the column, else scientific
    static std::string StringFromTW(double t, int width, int precision)
    {
        char b[96];
        snprintf(b, sizeof(b), "%.*f", precision, t);
        if ((int)std::strlen(b) <= width) {
            snprintf(b, sizeof(b), "%*.*f", width, precision, t);
            return b;
        }
        const int need = t < 0.0 ? 6 : 5;
        if (width < need)
            return std::string((size_t)width, '#');
        int prec1 = width - need;
        if (prec1 > 0)
            prec1--;
        snprintf(b, sizeof(b), "%*.*e", width, std::min(precision, prec1), t);
        return b;
    }
    static double removeNegativeZero(double t, int precision)
    {
        if (t < 0.0 || (t == 0.0 && std::signbit(t)))
            return std::fabs(std::floor(t * std::pow(10.0, precision) + 0.5)) > 0.0 ? t : 0.0;
        return t;
    }
    static std::string Fixed(double v, int width, int precision)
    {
        char b[96];
        snprintf(b, sizeof(b), "%*.*f", width, precision, v);
        return b;
    }
    // a number that may blow up in a questionable adjustment: column-safe notation then (PRN:2216-2223)
    std::string Num(double v, int width, int precision, bool safe) const { return safe ? StringFromTW(v, width, precision) : Fixed(v, width, precision); }

    // "ddd mm ss.ssss" / symbols / ddd.mmssssss / decimal degrees of an angular measurement (FormatAngularMeasurement PRN:2184-2260)
    static std::string dms_fields(double rad, int sec_precision, char* sign, long long* d, long long* mi, long long* s_int, long long* s_frac)
    {
        const double deg = std::fabs(rad) * 180.0 / 3.14159265358979323846;
        long long scale = 1;
        for (int k = 0; k < sec_precision; ++k)
            scale *= 10;
        const long long units = std::llround(deg * 3600.0 * (double)scale);   // carries are exact in integer arithmetic
        *d = units / (3600LL * scale);
        const long long rem = units % (3600LL * scale);
        *mi = rem / (60LL * scale);
        const long long sec = rem % (60LL * scale);
        *s_int = sec / scale;
        *s_frac = sec % scale;
        *sign = rad < 0 ? '-' : 0;
        return std::string();
    }
    std::string AngleString(double rad, int sec_precision, int angular_type, int dms_format) const
    {
        char b[96];
        if (angular_type == 1) {   // DDEG
            snprintf(b, sizeof(b), "%.*f", 4 + sec_precision, rad * 180.0 / 3.14159265358979323846);
            return b;
        }
        char sign;
        long long d, mi, si, sf;
        dms_fields(rad, sec_precision, &sign, &d, &mi, &si, &sf);
        const std::string sg = sign ? "-" : "";
        char frac[32] = "";
        if (sec_precision > 0)
            snprintf(frac, sizeof(frac), "%0*lld", sec_precision, sf);
        switch (dms_format) {
        case 1:   // ddd°mm'ss.sss" (Latin-1 symbols as the reference writes them)
            snprintf(b, sizeof(b), "%s%lld\260%02lld\222%02lld%s%s\224", sg.c_str(), d, mi, si, sec_precision > 0 ? "." : "", frac);
            break;
        case 2:   // ddd.mmssssss
            snprintf(b, sizeof(b), "%s%lld.%02lld%02lld%s", sg.c_str(), d, mi, si, frac);
            break;
        default:  // ddd mm ss.ssss
            snprintf(b, sizeof(b), "%s%lld %02lld %02lld%s%s", sg.c_str(), d, mi, si, sec_precision > 0 ? "." : "", frac);
        }
        return b;
    }

    struct MsrRow {           // one printed row of the table, in the units of its frame
        char type, cardinal;
        bool angular, ignore;
        const char *s1, *s2, *s3;
        double measured, adjusted, corr, var, adj_prec, res_prec, nstat, tstat, pelzer, pre_adj_corr;
        bool pre_adj_corr_linear;   // the H row of a geographic Y cluster prints its N value in metres
        bool show_type;
        int64_t rec;                // binary record the row belongs to (database ids)
    };

    void PrintMsrRow(std::ostream& os, const MsrRow& r, int mode /*0 adjusted, 1 computed / ignored*/) const
    {
        const double crit = stats_.critical_value;
        const bool safe = std::fabs(r.nstat) > crit * 4.0;
        const double SEC = 3.14159265358979323846 / 180.0 / 3600.0, DEG = 3.14159265358979323846 / 180.0;
        const int pa = a_.precision_seconds_msr, pl = a_.precision_metres_msr;
        char head[80];
        snprintf(head, sizeof(head), "%-2s%-20s%-20s%-20s%-3s%-2c", r.show_type ? std::string(1, r.type).c_str() : "", r.s1, r.s2, r.s3,
                 r.ignore ? "*" : " ", r.cardinal);
        os << head;
        if (r.angular) {
            const double unit = a_.angular_type_msr == 1 ? DEG : SEC;
            os << std::setw(19) << std::right << AngleString(r.measured, pa, a_.angular_type_msr, a_.dms_format_msr)
               << std::setw(19) << std::right << AngleString(r.adjusted, pa, a_.angular_type_msr, a_.dms_format_msr)
               << Num(removeNegativeZero(r.corr / unit, pa), 12, pa, safe) << Num(std::sqrt(r.var) / unit, 13, pa, safe);
            if (mode == 0)
                os << Num(std::sqrt(std::fabs(r.adj_prec)) / unit, 13, pa, safe) << Num(std::sqrt(r.res_prec) / unit, 13, pa, safe);
        } else {
            os << Fixed(r.measured, 19, pl) << Fixed(r.adjusted, 19, pl) << Num(removeNegativeZero(r.corr, pl), 12, pl, safe)
               << Num(std::sqrt(r.var), 13, pl, safe);
            if (mode == 0)
                os << Num(std::sqrt(std::fabs(r.adj_prec)), 13, pl, safe) << Num(std::sqrt(r.res_prec), 13, pl, safe);
        }
        if (mode == 0) {
            os << Num(removeNegativeZero(r.nstat, 2), 11, 2, safe);
            if (a_.adj_msr_tstat)
                os << Num(removeNegativeZero(r.tstat, 2), 11, 2, safe);
            os << Fixed(r.pelzer, 12, 2);
        }
        // pre-adjustment correction (PrintMeasurementCorrection PRN:2435-2486): seconds for the angular types, else metres
        if (std::strchr("ABDIJKPQVZ", r.type))
            os << Fixed(removeNegativeZero(r.pre_adj_corr / SEC, pa), 14, pa);
        else if (r.type == 'Y')
            os << Fixed(r.pre_adj_corr_linear ? removeNegativeZero(r.pre_adj_corr, pl) : 0.0, 14, (r.angular || r.cardinal == 'h') ? pa : pl);
        else
            os << Fixed(removeNegativeZero(r.pre_adj_corr, pa), 14, pa);
        if (mode == 0)
            os << std::setw(12) << std::right << (std::fabs(r.nstat) > crit ? "*" : " ");
        if (a_.database_ids && r.rec >= 0)
            PrintMeasurementDatabaseID(os, (size_t)r.rec);
        os << "\n";
    }

    // measurement id, and for D G X Y the cluster id, of a record (PrintMeasurementDatabaseID PRN:239-263)
    void PrintMeasurementDatabaseID(std::ostream& os, size_t rec) const
    {
        if (rec >= dbid_.size())
            return;
        const DbId& d = dbid_[rec];
        if (d.msr_set)
            os << std::setw(10) << std::right << d.msr_id;
        else
            os << std::setw(10) << " ";
        if (std::strchr("DGXY", msr_[rec].measType)) {
            if (d.cls_set)
                os << std::setw(10) << std::right << d.cluster_id;
            else
                os << std::setw(10) << " ";
        }
    }
    // <net>.dbid of dnaimport (LoadDatabaseId ADJ:2211-2276): u32 count, then per binary measurement record u32 measurement
    // id, u32 cluster id, u16 / u16 "is set" flags
    struct DbId {
        uint32_t msr_id, cluster_id;
        bool msr_set, cls_set;
    };
    void LoadDatabaseId()
    {
        const std::string file = a_.output_folder + "/" + a_.network_name + ".dbid";
        std::ifstream in(file, std::ios::binary);
        uint32_t count = 0;
        if (!in || !in.read(reinterpret_cast<char*>(&count), sizeof(count)))
            SignalExceptionAdjustment("LoadDatabaseId(): could not open " + file + " (written by dnaimport; needed for --output-database-ids)");
        dbid_.resize(count);
        for (uint32_t r = 0; r < count; ++r) {
            uint16_t a = 0, b = 0;
            in.read(reinterpret_cast<char*>(&dbid_[r].msr_id), 4);
            in.read(reinterpret_cast<char*>(&dbid_[r].cluster_id), 4);
            in.read(reinterpret_cast<char*>(&a), 2);
            in.read(reinterpret_cast<char*>(&b), 2);
            dbid_[r].msr_set = a != 0;
            dbid_[r].cls_set = b != 0;
        }
        if (!in)
            SignalExceptionAdjustment("LoadDatabaseId(): " + file + " is truncated");
        a_.output_msr_blocks = false;   // ids go with one contiguous list in the original order (ADJ:2222-2226)
    }

    MsrRow ScalarRow(const dna_msr_t& m, char cardinal, double var) const
    {
        MsrRow r{};
        r.type = m.measType;
        r.cardinal = cardinal;
        r.angular = std::strchr("ABDKVZIJPQ", m.measType) != nullptr;
        r.ignore = m.ignore != 0;
        r.s1 = stn_[m.station1].stationName;
        r.s2 = r.s3 = "";
        r.measured = m.preAdjMeas;
        r.adjusted = m.measAdj;
        r.corr = m.measCorr;
        r.var = var;
        r.adj_prec = m.measAdjPrec;
        r.res_prec = m.residualPrec;
        r.nstat = m.NStat;
        r.tstat = m.TStat;
        r.pelzer = m.PelzerRel;
        r.pre_adj_corr = m.preAdjCorr;
        r.pre_adj_corr_linear = false;
        r.show_type = true;
        r.rec = (&m >= msr_.data() && &m < msr_.data() + msr_.size()) ? (int64_t)(&m - msr_.data()) : -1;
        return r;
    }

    void PrintMsrTableHeader(std::ostream& os, const std::string& heading, int mode) const
    {
        os << "\n" << heading << "\n------------------------------------------\n\n";
        char buf[512];
        snprintf(buf, sizeof(buf), "%-2s%-20s%-20s%-20s%-3s%-2s%19s%19s%12s%13s", "M", "Station 1", "Station 2", "Station 3", "*", "C", "Measured",
                 mode == 0 ? "Adjusted" : "Computed", mode == 0 ? "Correction" : "Difference", "Meas. SD");
        os << buf;
        size_t width = 2 + 60 + 3 + 3 + 19 + 19 + 12 + 13;
        if (mode == 0) {
            os << std::setw(13) << std::right << "Adj. SD" << std::setw(13) << "Corr. SD" << std::setw(11) << "N-stat";
            width += 13 + 13 + 11;
            if (a_.adj_msr_tstat) {
                os << std::setw(11) << "T-stat";
                width += 11;
            }
            os << std::setw(12) << "Pelzer Rel";
            width += 12;
        }
        os << std::setw(14) << std::right << "Pre Adj Corr";
        width += 14;
        if (mode == 0) {
            os << std::setw(12) << "Outlier?";
            width += 12;
        }
        if (a_.database_ids) {
            os << std::setw(10) << "Meas. ID" << std::setw(10) << "Clust. ID";
            width += 20;
        }
        os << "\n" << std::string(width, '-') << "\n";
    }

    void PrintAdjMeasurements(std::ostream& os, const std::vector<int32_t>* rec_block, int32_t block, const std::string& heading = "Adjusted Measurements") const
    {
        PrintMsrTableHeader(os, heading, 0);
        std::vector<uint32_t> list = CollectMeasurements(rec_block, block, false);
        SortMeasurements(list);
        PrintMeasurementRecords(os, list, 0);
        os << "\n";
    }

    // "Ignored Measurements (a-posteriori)" (PrintIgnoredAdjMeasurements PRN:1784-1923): measured, computed from the
    // adjusted coordinates, difference, measurement SD and pre-adjustment correction of every ignored measurement
    void PrintIgnoredAdjMeasurements(std::ostream& os)
    {
        if (ctx_)   // report mode prints what the last adjustment left in the records
            check(gadj_update_ignored_measurements(ctx_));
        PrintMsrTableHeader(os, "Ignored Measurements (a-posteriori)", 1);
        PrintMeasurementRecords(os, CollectMeasurements(nullptr, -1, true), 1);
        os << "\n\n";
    }

    void PrintMeasurementRecords(std::ostream& os, const std::vector<uint32_t>& list, int mode) const
    {
        for (uint32_t first : list) {
            const dna_msr_t& m = msr_[first];
            switch (m.measType) {
            case 'G': case 'X': case 'Y':
                PrintMeasurements_GXY(os, first, mode);
                break;
            case 'D':
                PrintMeasurements_D(os, first, mode);
                break;
            default: {
                MsrRow r = ScalarRow(m, ' ', m.term2);
                if (m.measurementStations >= 2)
                    r.s2 = stn_[m.station2].stationName;
                if (m.measurementStations >= 3 && m.measType == 'A')
                    r.s3 = stn_[m.station3].stationName;
                PrintMsrRow(os, r, mode);
            }
            }
        }
    }

    // a direction set: one heading row (instrument, reference object, number of angles), then the derived angles, each
    // against its target (PrintAdjMeasurements_D PRN:917-975); measured / adjusted are the direction itself and the
    // direction plus the angle's correction (PRN:2309-2316), the precision that of the derived angle (scale2)
    void PrintMeasurements_D(std::ostream& os, uint32_t first, int mode) const
    {
        const dna_msr_t& ro = msr_[first];
        const uint32_t angles = ro.vectorCount2 > 0 ? ro.vectorCount2 - 1 : 0;
        char head[96];
        snprintf(head, sizeof(head), "%-2c%-20s%-20s%-20s%-3s%-2u", 'D', stn_[ro.station1].stationName, stn_[ro.station2].stationName, "",
                 ro.ignore ? "*" : " ", angles);
        os << head;
        if (a_.database_ids) {
            os << std::string(19 + 19 + 12 + 13 + (mode == 0 ? 13 + 13 + 11 + (a_.adj_msr_tstat ? 11 : 0) + 12 : 0) + 14 + (mode == 0 ? 12 : 0), ' ');
            PrintMeasurementDatabaseID(os, first);
        }
        os << "\n";
        uint32_t printed = 0;
        for (size_t j = first + 1; j < first + std::max<uint32_t>(1u, ro.vectorCount1) && j < msr_.size() && printed < angles; ++j) {
            const dna_msr_t& d = msr_[j];
            if (d.ignore && !ro.ignore)
                continue;
            MsrRow r = ScalarRow(d, ' ', d.scale2);
            r.show_type = false;
            r.ignore = false;
            r.s1 = r.s2 = "";
            r.s3 = stn_[d.station2].stationName;
            r.measured = d.term1;
            r.adjusted = d.term1 + d.measCorr;
            PrintMsrRow(os, r, mode);
            ++printed;
        }
    }

    // G baselines and X / Y clusters (PrintAdjMeasurements_GXY PRN:4072-4144): three rows per member
    void PrintMeasurements_GXY(std::ostream& os, uint32_t first, int mode) const
    {
        const dna_msr_t& c = msr_[first];
        const uint32_t count = std::max<uint32_t>(1u, c.vectorCount1);
        size_t j = first;
        for (uint32_t k = 0; k < count && j + 2 < msr_.size(); ++k) {
            const dna_msr_t* r = &msr_[j];
            if (c.measType == 'Y' && mode == 1 && std::strncmp(r->coordType, "LL", 2) == 0) {
                // an ignored cluster was never converted: its records still hold latitude, longitude, height
                const double var[3] = {r[0].term2, r[1].term3, r[2].term4};
                for (int q = 0; q < 3; ++q) {
                    MsrRow row = ScalarRow(r[q], q == 0 ? 'P' : q == 1 ? 'L' : (std::strncmp(r->coordType, "LLH", 3) == 0 ? 'H' : 'h'), var[q]);
                    row.angular = q < 2;
                    row.pre_adj_corr_linear = q == 2;
                    PrintMsrRow(os, row, mode);
                }
            } else if (c.measType == 'Y' && (r->station3 == DNA_LLH_TYPE || r->station3 == DNA_LLh_TYPE))
                PrintMeasurements_YLLH(os, j, mode);
            else if (a_.adj_gnss_units != 0 && c.measType != 'Y' && mode == 0)
                PrintAdjGNSSAlternateUnits(os, j);
            else {
                const double var[3] = {r[0].term2, r[1].term3, r[2].term4};
                for (int q = 0; q < 3; ++q) {
                    MsrRow row = ScalarRow(r[q], "XYZ"[q], var[q]);
                    row.s1 = stn_[r->station1].stationName;
                    row.s2 = c.measType == 'Y' ? "" : stn_[r->station2].stationName;
                    PrintMsrRow(os, row, mode);
                }
            }
            j += 3 + 3 * (size_t)r->vectorCount2;
        }
    }

    // A point of a Y cluster that was supplied as latitude / longitude / height is reported in that form
    // (PrintAdjMeasurements_YLLH PRN:2488-2660, ReduceYLLHMeasurementsforPrinting ADJ:9981-10046): the adjusted Cartesian
    // point goes back to geographic (orthometric height for LLH: minus the geoid separation), the corrections are taken
    // against the original values kept in preAdjMeas, and the variances of the measurement (its 3x3 Cartesian block) and
    // of the adjusted measurement (its three Cartesian variances) are propagated to geographic with the Jacobian at the
    // adjusted position; N-stat and Pelzer reliability are then recomputed in that frame.
    void PrintMeasurements_YLLH(std::ostream& os, size_t i, int mode) const
    {
        const dna_msr_t* r = &msr_[i];
        const dna_stn_t& st = stn_[r->station1];
        const gadj::Ellipsoid ell = Ellipsoid();
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
        const double sz = std::sqrt(stats_.sigma_zero);
        for (int q = 0; q < 3; ++q) {
            MsrRow row = ScalarRow(r[q], comp[q], var[q]);
            row.angular = q < 2;
            row.adjusted = adj[q];
            row.corr = adj[q] - r[q].preAdjMeas;
            row.adj_prec = adjp[q];
            row.res_prec = std::fabs(var[q] - adjp[q]);
            row.pelzer = std::sqrt(var[q]) / std::sqrt(row.res_prec);
            if (!(row.pelzer >= 0.0) || row.pelzer > 700.0)
                row.pelzer = 999.99;
            row.nstat = row.corr / std::sqrt(row.res_prec);
            row.tstat = sz > 1.0e-10 ? row.nstat / sz : 0.0;
            row.pre_adj_corr_linear = comp[q] == 'H';
            PrintMsrRow(os, row, mode);
        }
    }

    // --output-adj-gnss-units 1 | 2 | 3: a G / X baseline in the local frame at its first station — east north up;
    // azimuth, vertical angle, slope distance; or azimuth, slope distance, up (PrintAdjGNSSAlternateUnits PRN:4717-5047).
    // Variances of the measurement and of the adjusted measurement (Q11 + Q22 - Q12 - Q21) are rotated with the local
    // frame at the mid point of the line, then to polar with the Jacobian of (azimuth, elevation, distance); statistics
    // are recomputed per component (UpdateMsrRecordStats ADJ:8283-8290).
    void PrintAdjGNSSAlternateUnits(std::ostream& os, size_t i) const
    {
        const dna_msr_t* r = &msr_[i];
        const dna_stn_t &s1 = stn_[r->station1], &s2 = stn_[r->station2];
        double Vm[9] = {r[0].term2, r[1].term2, r[2].term2, r[1].term2, r[1].term3, r[2].term3, r[2].term2, r[2].term3, r[2].term4};
        double Va[9];
        BaselinePrecision(i, Va);
        const double meas[3] = {r[0].term1, r[1].term1, r[2].term1}, adjm[3] = {r[0].measAdj, r[1].measAdj, r[2].measAdj};
        double R1[9], Rm[9];
        local_rotation(s1.currentLatitude, s1.currentLongitude, R1);
        local_rotation(0.5 * (s1.currentLatitude + s2.currentLatitude), 0.5 * (s1.currentLongitude + s2.currentLongitude), Rm);
        double ml[3], al[3];
        for (int k = 0; k < 3; ++k) {   // cart -> local: R^T v
            ml[k] = R1[k] * meas[0] + R1[3 + k] * meas[1] + R1[6 + k] * meas[2];
            al[k] = R1[k] * adjm[0] + R1[3 + k] * adjm[1] + R1[6 + k] * adjm[2];
        }
        double Vl[9], Val[9];
        rotate_sym(Rm, Vm, Vl);
        rotate_sym(Rm, Va, Val);
        double measured[3], adjusted[3], var[3], adjp[3];
        char card[3];
        bool ang[3] = {false, false, false};
        if (a_.adj_gnss_units == 1) {
            for (int k = 0; k < 3; ++k) {
                measured[k] = ml[k];
                adjusted[k] = al[k];
                var[k] = Vl[4 * k];
                adjp[k] = Val[4 * k];
                card[k] = "enu"[k];
            }
        } else {
            const double az = direction_en(ml[0], ml[1]), el = std::atan2(ml[2], std::hypot(ml[0], ml[1]));
            const double dist = std::sqrt(ml[0] * ml[0] + ml[1] * ml[1] + ml[2] * ml[2]);
            const double azA = direction_en(al[0], al[1]), elA = std::atan2(al[2], std::hypot(al[0], al[1]));
            const double distA = std::sqrt(al[0] * al[0] + al[1] * al[1] + al[2] * al[2]);
            // Jacobian local -> polar (FormLocaltoPolarRotationMatrix MFN:482-504)
            const double ca = std::cos(az), sa = std::sin(az), ce = std::cos(el), se = std::sin(el);
            const double P[9] = {ca / dist, -sa / dist, 0.0, -sa * se / dist, -ca * se / dist, ce / dist, sa * ce, ca * ce, se};
            double vp[3], vap[3];
            for (int a = 0; a < 3; ++a) {
                vp[a] = vap[a] = 0.0;
                for (int x = 0; x < 3; ++x)
                    for (int y = 0; y < 3; ++y) {
                        vp[a] += P[3 * a + x] * Vl[3 * x + y] * P[3 * a + y];
                        vap[a] += P[3 * a + x] * Val[3 * x + y] * P[3 * a + y];
                    }
            }
            if (a_.adj_gnss_units == 2) {   // azimuth, vertical angle, slope distance
                const double m3[3] = {az, el, dist}, a3[3] = {azA, elA, distA};
                for (int k = 0; k < 3; ++k) {
                    measured[k] = m3[k];
                    adjusted[k] = a3[k];
                    var[k] = vp[k];
                    adjp[k] = vap[k];
                }
                card[0] = 'a', card[1] = 'v', card[2] = 's';
                ang[0] = ang[1] = true;
            } else {                        // azimuth, slope distance, up
                measured[0] = az, adjusted[0] = azA, var[0] = vp[0], adjp[0] = vap[0];
                measured[1] = dist, adjusted[1] = distA, var[1] = vp[2], adjp[1] = vap[2];
                measured[2] = ml[2], adjusted[2] = al[2], var[2] = Vl[8], adjp[2] = Val[8];
                card[0] = 'a', card[1] = 's', card[2] = 'u';
                ang[0] = true;
            }
        }
        const double sz = std::sqrt(stats_.sigma_zero);
        for (int q = 0; q < 3; ++q) {
            MsrRow row = ScalarRow(r[q], card[q], var[q]);
            row.s1 = s1.stationName;
            row.s2 = s2.stationName;
            row.angular = ang[q];
            row.measured = measured[q];
            row.adjusted = adjusted[q];
            row.corr = adjusted[q] - measured[q];
            row.adj_prec = adjp[q];
            row.res_prec = var[q] - adjp[q];
            row.pelzer = std::sqrt(var[q]) / std::sqrt(row.res_prec);
            if (!(row.pelzer >= 0.0) || row.pelzer > 700.0)
                row.pelzer = 999.99;
            row.nstat = row.corr / std::sqrt(row.res_prec);
            row.tstat = sz > 1.0e-10 ? row.nstat / sz : 0.0;
            PrintMsrRow(os, row, 0);
        }
    }

    // columns of R: east, north, up in Cartesian components (local -> cart)
    static void local_rotation(double lat, double lon, double* R)
    {
        const double sl = std::sin(lat), cl = std::cos(lat), so = std::sin(lon), co = std::cos(lon);
        const double M[9] = {-so, -sl * co, cl * co, co, -sl * so, cl * so, 0.0, cl, sl};
        std::memcpy(R, M, sizeof(M));
    }
    static void rotate_sym(const double* R, const double* V, double* out)   // R^T V R
    {
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                double s = 0.0;
                for (int x = 0; x < 3; ++x)
                    for (int y = 0; y < 3; ++y)
                        s += R[3 * x + a] * V[3 * x + y] * R[3 * y + b];
                out[3 * a + b] = s;
            }
    }
    gadj::Ellipsoid Ellipsoid() const
    {
        gadj_opts o;
        gadj_default_opts(&o);
        return gadj::make_ellipsoid(o.semi_major, o.inv_flattening);
    }
    void check_const(int rc) const
    {
        if (rc)
            throw std::runtime_error(gadj_last_error(ctx_));
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

    // Redfearn's formulae, geographic -> UTM / MGA grid (GeoToGrid GEO:365-432; K0 0.9996, false origin 500 000 / 10 000 000,
    // 6 degree zones, zone 0 central meridian -183)
    static void GeoToGrid(const gadj::Ellipsoid& ell, double lat, double lon, double* easting, double* northing, double* zone)
    {
        const double PI = 3.14159265358979323846, K0 = 0.9996;
        *zone = std::floor((lon * 180.0 / PI + 186.0) / 6.0);
        const double w = lon - (*zone * 6.0 - 183.0) * PI / 180.0;
        const double e2 = ell.e2, e4 = e2 * e2, e6 = e4 * e2;
        const double s = std::sin(lat), c = std::cos(lat), t = std::tan(lat), t2 = t * t, t4 = t2 * t2, t6 = t4 * t2;
        const double nu = ell.a / std::sqrt(1.0 - e2 * s * s), rho = ell.a * (1.0 - e2) / std::pow(1.0 - e2 * s * s, 1.5), psi = nu / rho;
        const double A0 = 1.0 - e2 / 4.0 - 3.0 * e4 / 64.0 - 5.0 * e6 / 256.0, A2 = 3.0 / 8.0 * (e2 + e4 / 4.0 + 15.0 * e6 / 128.0);
        const double A4 = 15.0 / 256.0 * (e4 + 3.0 * e6 / 4.0), A6 = 35.0 * e6 / 3072.0;
        const double m = ell.a * (A0 * lat - A2 * std::sin(2 * lat) + A4 * std::sin(4 * lat) - A6 * std::sin(6 * lat));
        const double w2 = w * w, w4 = w2 * w2, w6 = w4 * w2, w8 = w4 * w4, c2 = c * c;
        const double E1 = w2 / 6.0 * c2 * (psi - t2);
        const double E2 = w4 / 120.0 * c2 * c2 * (4.0 * psi * psi * psi * (1.0 - 6.0 * t2) + psi * psi * (1.0 + 8.0 * t2) - psi * 2.0 * t2 + t4);
        const double E3 = w6 / 5040.0 * c2 * c2 * c2 * (61.0 - 479.0 * t2 + 179.0 * t4 - t6);
        *easting = K0 * nu * w * c * (1.0 + E1 + E2 + E3) + 500000.0;
        const double N1 = w2 / 2.0 * nu * s * c;
        const double N2 = w4 / 24.0 * nu * s * c * c2 * (4.0 * psi * psi + psi - t2);
        const double N3 = w6 / 720.0 * nu * s * c * c2 * c2 *
                          (8.0 * psi * psi * psi * psi * (11.0 - 24.0 * t2) - 28.0 * psi * psi * psi * (1.0 - 6.0 * t2) + psi * psi * (1.0 - 32.0 * t2) -
                           psi * 2.0 * t2 + t4);
        const double N4 = w8 / 40320.0 * nu * s * c * c2 * c2 * c2 * (1385.0 - 3111.0 * t2 + 543.0 * t4 - t6);
        *northing = K0 * (m + N1 + N2 + N3 + N4) + 10000000.0;
    }

    // the station coordinates the corrections are measured from (v_originalStations_; re-derived from the initial
    // coordinates of the station file when corrections are reported, PRN:3934-3950)
    void OriginalXYZ(size_t i, double* xyz) const
    {
        if (a_.stn_corrections || a_.output_corrections) {
            const dna_stn_t& s = stn_[i];
            double h = s.initialHeight;
            if (s.suppliedHeightRefFrame == 0)   // ORTHOMETRIC_type_i
                h += s.geoidSep;
            gadj::geo_to_cart(Ellipsoid(), s.initialLatitude, s.initialLongitude, h, xyz);
            return;
        }
        std::memcpy(xyz, &apriori_xyz_[3 * i], 3 * sizeof(double));
    }

    std::vector<uint32_t> StationOrder(const std::vector<uint32_t>* subset) const
    {
        std::vector<uint32_t> list;
        if (subset)
            list = *subset;
        else {
            list.resize(stn_.size());
            for (size_t i = 0; i < list.size(); ++i)
                list[i] = (uint32_t)i;
        }
        if (a_.sort_stn_orig_order)   // --sort-stn-orig-order: the order of the imported station file (CompareStnFileOrder)
            std::stable_sort(list.begin(), list.end(), [&](uint32_t a, uint32_t b) { return stn_[a].fileOrder < stn_[b].fileOrder; });
        return list;
    }

    void PrintAdjStations(std::ostream& os, const std::vector<uint32_t>* subset, const std::string& heading = "Adjusted Coordinates") const
    {   // PrintAdjStation (PRN:3917-4070): the coordinate types of --stn-coord-types + SD(e,n,up) = sqrt diag(R^T Q R),
        // geoid uncertainty added to up; optional corrections (e, n, up) from the original coordinates
        os << "\n" << heading << "\n------------------------------------------\n\n";
        const std::string& types = a_.stn_coord_types;
        const int pl = a_.precision_metres_stn, pa = a_.precision_seconds_stn;
        auto width_of = [](char c) { return c == 'P' || c == 'E' ? 14 : c == 'L' || c == 'N' ? 15 : c == 'H' || c == 'h' ? 11 : c == 'z' ? 8 : 15; };
        auto name_of = [](char c) -> const char* {
            switch (c) {
            case 'P': return "Latitude";
            case 'L': return "Longitude";
            case 'H': return "H(Ortho)";
            case 'h': return "h(Ellipse)";
            case 'E': return "Easting";
            case 'N': return "Northing";
            case 'z': return "Zone";
            case 'X': return "X";
            case 'Y': return "Y";
            case 'Z': return "Z";
            }
            return "";
        };
        os << std::left << std::setw(20) << "Station" << std::setw(5) << "Const";
        size_t width = 25;
        for (char c : types) {
            if (!std::strchr("PLHhENzXYZ", c))
                continue;
            os << std::right << std::setw(width_of(c)) << name_of(c);
            width += width_of(c);
        }
        os << "  " << std::right << std::setw(10) << "SD(e)" << std::setw(10) << "SD(n)" << std::setw(10) << "SD(up)";
        width += 2 + 30 + 2 + 56;
        if (a_.stn_corrections) {
            os << "  " << std::setw(11) << "Corr(e)" << std::setw(11) << "Corr(n)" << std::setw(11) << "Corr(up)";
            width += 2 + 33;
        }
        os << "  " << std::left << "Description" << "\n" << std::string(width, '-') << "\n";
        const bool grid = types.find_first_of("ENz") != std::string::npos;
        const gadj::Ellipsoid ell = Ellipsoid();
        for (uint32_t i : StationOrder(subset)) {
            const dna_stn_t& s = stn_[i];
            const double* q = &vcv_[9 * (size_t)i];
            const double lat = s.currentLatitude, lon = s.currentLongitude, h = s.currentHeight;
            double E = 0, N = 0, zone = -1;
            if (grid)
                GeoToGrid(ell, lat, lon, &E, &N, &zone);
            char cst[4] = {s.stationConst[0], s.stationConst[1], s.stationConst[2], 0};
            os << std::left << std::setw(20) << s.stationName << std::setw(5) << cst << std::right;
            for (char c : types) {
                switch (c) {
                case 'P':
                    os << std::setw(14) << (a_.angular_type_stn == 1 ? Fixed(lat * 180.0 / 3.14159265358979323846, 0, 4 + pa) : hp_dms(lat, 4 + pa));
                    break;
                case 'L':
                    os << std::setw(15) << (a_.angular_type_stn == 1 ? Fixed(lon * 180.0 / 3.14159265358979323846, 0, 4 + pa) : hp_dms(lon, 4 + pa));
                    break;
                case 'E': os << Fixed(E, 14, pl); break;
                case 'N': os << Fixed(N, 15, pl); break;
                case 'z': os << Fixed(zone, 8, 0); break;
                case 'H': os << Fixed(h - (double)s.geoidSep, 11, pl); break;
                case 'h': os << Fixed(h, 11, pl); break;
                case 'X': os << Fixed(est_[3 * (size_t)i], 15, pl); break;
                case 'Y': os << Fixed(est_[3 * (size_t)i + 1], 15, pl); break;
                case 'Z': os << Fixed(est_[3 * (size_t)i + 2], 15, pl); break;
                }
            }
            double R[9], ql[9];
            local_rotation(lat, lon, R);
            rotate_sym(R, q, ql);
            ql[8] += (double)s.geoidSepUnc * s.geoidSepUnc;
            os << "  ";
            for (int k = 0; k < 3; ++k)
                os << Fixed(std::sqrt(std::fabs(ql[4 * k])), 10, pl);
            if (a_.stn_corrections) {
                double o[3];
                OriginalXYZ(i, o);
                const double d[3] = {est_[3 * (size_t)i] - o[0], est_[3 * (size_t)i + 1] - o[1], est_[3 * (size_t)i + 2] - o[2]};
                os << "  ";
                for (int k = 0; k < 3; ++k)
                    os << Fixed(removeNegativeZero(R[k] * d[0] + R[3 + k] * d[1] + R[6 + k] * d[2], pl), 11, pl);
            }
            os << "  " << s.description << "\n";
        }
        os << "\n";
    }

    // ---- "Measurements to Station" table (PrintMeasurementsToStation PRN:720-789): per station, the number of
    // non-ignored measurements of every type it takes part in (a cluster or direction set counts once per station)
    void PrintMeasurementsToStation(std::ostream& os) const
    {
        static const char kTypes[] = "ABCDEGHIJKLMPQRSVXYZ";
        std::vector<std::array<uint32_t, 20>> tally(stn_.size());
        for (auto& t : tally)
            t.fill(0);
        std::vector<uint32_t> touched;
        for (size_t i = 0; i < msr_.size();) {
            const size_t span = MeasurementSpan(i);
            const dna_msr_t& m = msr_[i];
            const char* p = std::strchr(kTypes, m.measType);
            if (!m.ignore && p && m.measType) {
                touched.clear();
                for (size_t j = i; j < i + span && j < msr_.size(); ++j) {
                    const dna_msr_t& r = msr_[j];
                    if (r.ignore || (std::strchr("GXY", r.measType) && r.measStart != 0))
                        continue;
                    touched.push_back(r.station1);
                    if (r.measurementStations >= 2 && r.measType != 'Y')
                        touched.push_back(r.station2);
                    if (r.measurementStations >= 3 && r.measType == 'A')
                        touched.push_back(r.station3);
                }
                std::sort(touched.begin(), touched.end());
                touched.erase(std::unique(touched.begin(), touched.end()), touched.end());
                for (uint32_t sidx : touched)
                    if (sidx < tally.size())
                        tally[sidx][p - kTypes]++;
            }
            i += span;
        }
        auto total_of = [&](uint32_t sidx) {
            uint32_t t = 0;
            for (uint32_t v : tally[sidx])
                t += v;
            return t;
        };
        auto line = [&]() { os << std::string(20 + 8 * 20 + 11, '-') << "\n"; };
        os << "\nMeasurements to Station \n------------------------------------------\n\n" << std::left << std::setw(20) << "Station";
        for (const char* c = kTypes; *c; ++c)
            os << std::right << std::setw(8) << *c;
        os << std::setw(11) << "Total" << "\n";
        line();
        std::vector<uint32_t> order(stn_.size());
        for (size_t i = 0; i < order.size(); ++i)
            order[i] = (uint32_t)i;
        switch (a_.sort_msr_to_stn) {   // orig_stn_sort_ui 0, name 1, count ascending 2, count descending 3
        case 0: std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return stn_[a].fileOrder < stn_[b].fileOrder; }); break;
        case 2: std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return total_of(a) < total_of(b); }); break;
        case 3: std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return total_of(a) > total_of(b); }); break;
        default: std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return stn_[a].nameOrder < stn_[b].nameOrder; });
        }
        auto row = [&](const char* name, const std::array<uint32_t, 20>& t) {
            os << std::left << std::setw(20) << name << std::right;
            uint32_t total = 0;
            for (uint32_t v : t) {
                if (v)
                    os << std::setw(8) << v;
                else
                    os << std::setw(8) << " ";
                total += v;
            }
            os << std::setw(11) << total << "\n";
        };
        std::array<uint32_t, 20> totals;
        totals.fill(0);
        for (uint32_t sidx : order) {
            row(stn_[sidx].stationName, tally[sidx]);
            for (int k = 0; k < 20; ++k)
                totals[k] += tally[sidx][k];
        }
        line();
        row("Totals", totals);
        os << "\n\n";
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
    std::vector<DbId> dbid_;
    std::vector<double> corrPrev_;
    std::vector<uint32_t> stnOscCount_;
    std::map<uint32_t, OscillationRecord> oscHistory_;
    std::vector<std::string> iter_pre_, iter_post_;   // per-iteration report text (--output-iter-*)
    std::vector<double> est_, vcv_, raw_vcv_, apriori_llh_, apriori_xyz_;
    std::vector<uint32_t> pam_rec_;     // first records of the G / X baselines, ascending
    std::vector<double> pam_;           // 6 per baseline: upper triangle of the variance of the adjusted baseline
    bool report_mode_ = false;
    uint32_t last_iterations_ = 0;
    double maxCorr_ = 0, total_ms_ = 0, chiUpper_ = 0, chiLower_ = 0;
    int passFail_ = 0;
    ADJUST_STATUS adjustStatus_ = ADJUST_SUCCESS;
};

}  // namespace dynadjust_b200
