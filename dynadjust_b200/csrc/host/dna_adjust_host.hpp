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
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <algorithm>
#include <array>
#include <atomic>
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
#include "gpu_group.hpp"

namespace dynadjust_b200 {

enum ADJUST_MODE { SimultaneousMode = 0, PhasedMode = 1, Phased_Block_1Mode = 2 };
enum ADJUST_STATUS { ADJUST_SUCCESS = 0, ADJUST_MAX_ITERATIONS_EXCEEDED = 2, ADJUST_EXCEPTION_RAISED = 4, ADJUST_CANCELLED = 5 };

struct adjust_settings {            // the fields of project_settings.a / .g / .o the solve path reads
    std::string network_name;
    std::string input_folder = ".", output_folder = ".";
    int adjust_mode = SimultaneousMode;
    int gpus = 1, first_device = 0;              // --gpus N: the adjustment sharded over N GPUs (one host thread per GPU)
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
        group_.stop();
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
        ComputeStationValidity();
        LoadAssociatedStationList(base + ".asl");
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
        o.device = a_.first_device;
        if (a_.gpus > 1 && o.ordering == GADJ_ORDER_DENSE)
            SignalExceptionAdjustment("--gpus: a single dense front cannot be sharded; drop the full-matrix exports or run on one GPU");
        if (gadj_create(&o, &ctx_))
            SignalExceptionAdjustment(gadj_last_error(nullptr));
        if (a_.gpus > 1)
            group_.start(a_.gpus, a_.first_device, o, stn_, msr_);   // worker ranks: own contexts, own copies of the records
        const int reduced = bms_meta_.reduced ? 1 : 0;
        all_ranks([&](gadj_ctx* c, int rank) -> int {
            if (a_.gpus > 1 && gadj_mg_init(c, rank, a_.gpus))
                return 1;
            return gadj_set_stations(c, rank ? group_.stn(rank) : stn_.data(), (uint32_t)stn_.size()) ||
                   gadj_set_measurements(c, rank ? group_.msr(rank) : msr_.data(), msr_.size()) ||
                   gadj_set_measurements_reduced(c, reduced);   // isFirstTimeAdjustment_ (ADJ:296)
        });
        if (a_.adjust_mode != SimultaneousMode) {
            if (a_.seg_file.empty()) {
                // a default .seg file older than station / measurement files that dnaimport wrote since describes another
                // network (WRAP:1239-1265); files updated by an adjustment keep their segmentation
                namespace fs = std::filesystem;
                auto imported = [](const dnafiles::BinaryMeta& m) {
                    std::string by(m.modifiedBy, strnlen(m.modifiedBy, sizeof(m.modifiedBy)));
                    for (char& ch : by)
                        ch = (char)std::tolower((unsigned char)ch);
                    return by == "import" || by == "dnaimport" || by == "dnainterop.dll" || by == "libdnaimport.so";
                };
                std::error_code ec1, ec2, ec3;
                const auto t_seg = fs::last_write_time(base + ".seg", ec1), t_bst = fs::last_write_time(bst_file_, ec2),
                           t_bms = fs::last_write_time(bms_file_, ec3);
                if (!ec1 && !ec2 && !ec3 && ((imported(bst_meta_) && t_seg < t_bst) || (imported(bms_meta_) && t_seg < t_bms)))
                    SignalExceptionAdjustment("The raw stations and measurements have been imported after\n  the segmentation file was created.\n"
                                              "  Run 'segment " + a_.network_name + " [options]' to re-create the segmentation file, or re-run\n"
                                              "  adjust using the --seg-file option if the file " + a_.network_name + ".seg must\n  be used.");
            }
            dnafiles::load_seg(a_.seg_file.empty() ? base + ".seg" : in_folder(a_.seg_file), seg_);
            std::vector<uint32_t> off{0}, isl;
            for (auto& b : seg_.isl) {
                isl.insert(isl.end(), b.begin(), b.end());
                off.push_back((uint32_t)isl.size());
            }
            all_ranks([&](gadj_ctx* c, int) { return gadj_set_blocks(c, (uint32_t)seg_.isl.size(), off.data(), isl.data()); });
        }
        all_ranks([&](gadj_ctx* c, int) { return gadj_prepare(c); });
        if (a_.gpus > 1) {
            // the ranks' buffer handles, once: from here on the GPUs talk to each other over NVLink
            std::vector<gadj_peer_info> peers(a_.gpus);
            all_ranks([&](gadj_ctx* c, int rank) { return gadj_mg_export(c, &peers[rank]); });
            all_ranks([&](gadj_ctx* c, int) { return gadj_mg_connect(c, peers.data()); });
        }
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
            all_ranks([&](gadj_ctx* c, int rank) {
                gadj_iter_result rr;
                return gadj_iterate(c, i == 0 ? GADJ_ITER_NORMALS : 0, rank ? &rr : &r);
            });
            r.ms_inverse = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - ti).count();  // wall
            iterations_.push_back(r);
            maxCorr_ = r.max_corr;
            UpdateIterationDiagnostics();
            iter_pre_.push_back(pre.str());
            iter_post_.push_back(std::string());
            if (std::fabs(r.max_corr) <= a_.iteration_threshold)
                break;
            if (cancel_ && cancel_->load()) {   // SIGINT: CancelAdjustment(), polled once per iteration (WRAP:67-71, ADJ:2432)
                adjustStatus_ = ADJUST_CANCELLED;
                break;
            }
            if (iter_reports && i + 1 < a_.max_iterations) {
                // --output-iter-adj-stat / -msr / -stn: statistics, adjusted measurements and stations of an iteration that
                // is followed by another (ADJ:2483-2502); needs the rigorous variances of this iteration
                all_ranks([&](gadj_ctx* c, int) { return gadj_form_inverse(c); });
                GenerateStatistics();
                if (a_.iter_adj_stat)
                    PrintStatisticsSummary(post, false);
                if (a_.iter_adj_msr) {
                    if (a_.adj_gnss_units != 0)
                        ComputeBaselinePrecisions();
                    PrintAdjMeasurements(post, nullptr, -1);
                }
                if (a_.iter_adj_stn)
                    PrintAdjStations(post, nullptr);
                iter_post_.back() = post.str();
            }
        }
        if (adjustStatus_ != ADJUST_CANCELLED && iterations_.size() == a_.max_iterations && std::fabs(maxCorr_) > a_.iteration_threshold)
            adjustStatus_ = ADJUST_MAX_ITERATIONS_EXCEEDED;   // ADJ:2523-2525
        if (adjustStatus_ == ADJUST_SUCCESS)
            all_ranks([&](gadj_ctx* c, int) { return gadj_form_inverse(c); });   // rigorous variances (v_rigorousVariances_)
        total_ms_ = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        return adjustStatus_;
    }

    // ---- GenerateStatistics (ADJ:6802) ------------------------------------------------------------------------
    void GenerateStatistics()
    {
        // also refreshes the records: adjusted lat/lon/h and measurement statistics
        all_ranks([&](gadj_ctx* c, int rank) {
            gadj_stats st;
            return gadj_statistics(c, rank ? &st : &stats_, 1);
        });
        est_.resize(3 * stn_.size());
        vcv_.resize(9 * stn_.size());
        check(gadj_get_estimates(ctx_, est_.data()));
        check(gadj_get_station_vcvs(ctx_, vcv_.data()));
        raw_vcv_ = vcv_;                              // before type B uncertainties: adjusted-measurement precisions use these
        pam_rec_.clear();                             // precisions of adjusted baselines belong to the variances just fetched
        pam_.clear();
        ApplyTypeBUncertainties();
        ComputeTestStat();
        if (a_.adj_msr_tstat) {                       // Student's t = n-stat / sqrt(sigma zero) (UpdateMsrTstatistic ADJ:6914-7092)
            const double sz = std::sqrt(stats_.sigma_zero);
            for (dna_msr_t& m : msr_)
                if (!m.ignore && m.measStart <= 2)
                    m.TStat = std::fabs(sz) < 1.0e-10 ? 0.0 : m.NStat / sz;
        }
    }

    void SetCancelFlag(const std::atomic<bool>* flag) { cancel_ = flag; }

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

#include "dna_adjust_reportmode.inl"

#include "dna_adjust_json.inl"

#include "dna_adjust_diagnostics.inl"

#include "dna_adjust_exports.inl"

#include "dna_adjust_apu_cor.inl"

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
        // station name -> .bst index: the station map <net>.map that dnaimport wrote (map_file.cpp:44-98, looked up by the
        // reference at LDR:243-244) when it is there and covers the station file, else the names in the .bst records
        std::unordered_map<std::string, uint32_t> by_name;
        {
            const std::string map_path = a_.input_folder + "/" + a_.network_name + ".map";
            std::vector<std::pair<std::string, uint32_t>> smap;
            if (std::filesystem::exists(map_path)) {
                dnafiles::load_map(map_path, smap);
                bool usable = smap.size() == stn_.size();
                for (const auto& e : smap)
                    usable = usable && e.second < stn_.size();
                if (usable)
                    for (const auto& e : smap)
                        by_name.emplace(e.first, e.second);
            }
            if (by_name.empty())
                for (size_t i = 0; i < stn_.size(); ++i)
                    by_name.emplace(stn_[i].stationName, (uint32_t)i);
        }
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
    // a call every rank of a multi-GPU run makes together (it contains device-side barriers); one GPU: just the call
    void all_ranks(const std::function<int(gadj_ctx*, int)>& fn)
    {
        if (group_.size() <= 1) {
            check(fn(ctx_, 0));
            return;
        }
        const std::string e = group_.all(fn, ctx_);
        if (!e.empty())
            SignalExceptionAdjustment(e);
    }
    // dense variance matrix of a block, from the rank that holds it (rank 0 holds its own subtrees and the shared top fronts)
    void block_vcv(uint32_t b, uint32_t* n, uint32_t* stations, uint32_t cap, double* packed)
    {
        if (gadj_get_block_vcv(ctx_, b, n, stations, cap, packed) == 0)
            return;
        const std::string first = gadj_last_error(ctx_);
        for (int r = 1; r < group_.size(); ++r)
            if (group_.one(r, [&](gadj_ctx* c, int) { return gadj_get_block_vcv(c, b, n, stations, cap, packed); }).empty())
                return;
        SignalExceptionAdjustment(first);
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

#include "dna_adjust_tables.inl"

    adjust_settings a_;
    gadj_ctx* ctx_ = nullptr;
    GpuGroup group_;             // ranks 1..N-1 of a --gpus N run
    gadj_info info_{};
    gadj_stats stats_{};
    std::vector<dna_stn_t> stn_;
    std::vector<dna_msr_t> msr_;
    dnafiles::BinaryMeta bst_meta_, bms_meta_;
    dnafiles::Segmentation seg_;
    std::string bst_file_, bms_file_;
    std::vector<gadj_iter_result> iterations_;
    const std::atomic<bool>* cancel_ = nullptr;
    std::vector<uint8_t> valid_;        // station takes part in the adjustment (a measurement that is not ignored touches it)
    size_t asl_differs_ = 0;            // stations whose <net>.asl validity disagrees with the measurement list
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
