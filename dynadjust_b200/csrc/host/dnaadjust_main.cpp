// dnaadjust_main.cpp — `dnaadjust <network> [options]`: the reference's command line for the adjust step
// (dnaadjustwrapper.cpp:799-1467; option strings config/dnaoptions-interface.hpp:96-158), driving the B200 engine.
// Flags the solve path honours are parsed; flags that only select the reference's CPU execution strategy
// (--multi-thread, --staged-adjustment, --create-stage-files, --purge-stage-files, --max-threads) are accepted and
// ignored; unknown flags are a hard error, as in the reference (WRAP:1040-1046).
#include <cstdlib>
#include <atomic>
#include <cctype>
#include <csignal>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "dna_adjust_host.hpp"

using namespace dynadjust_b200;

namespace {

struct Flag {
    const char* name;
    bool takes_value;
};

const Flag kFlags[] = {
    {"simultaneous-adjustment", false}, {"phased-adjustment", false}, {"block1-phased", false},
    {"staged-adjustment", false}, {"multi-thread", false}, {"report-results", false},
    {"conf-interval", true}, {"iteration-threshold", true}, {"max-iterations", true}, {"constraints", true},
    {"free-stn-sd", true}, {"fixed-stn-sd", true}, {"scale-normals-to-unity", false},
    {"create-stage-files", false}, {"purge-stage-files", false}, {"stage-path", true}, {"max-threads", true},
    {"input-folder", true}, {"output-folder", true}, {"output-adj-msr", false}, {"output-pos-uncertainty", false},
    {"output-all-covariances", false}, {"output-corrections-file", false}, {"output-apu-vcv-units", true},
    {"hz-corr-threshold", true}, {"vt-corr-threshold", true}, {"output-stn-blocks", false}, {"output-msr-blocks", false},
    {"export-sinex-file", false}, {"type-b-sd-global", true}, {"type-b-sd-file", true}, {"network-name", true}, {"quiet", false}, {"verbose-level", true},
    {"no-binary-update", false}, {"help", false},
    {"sort-adj-msr-field", true}, {"output-adj-gnss-units", true}, {"output-tstat-adj-msr", false}, {"output-msr-to-stn", false},
    {"sort-msr-to-stn-field", true}, {"stn-corrections", false}, {"stn-coord-types", true}, {"sort-stn-orig-order", false},
    {"angular-stn-type", true}, {"angular-msr-type", true}, {"dms-msr-format", true}, {"precision-stn-linear", true},
    {"precision-stn-angular", true}, {"precision-msr-linear", true}, {"precision-msr-angular", true}, {"output-iter-adj-stn", false},
    {"output-iter-adj-stat", false}, {"output-iter-adj-msr", false}, {"output-iter-cmp-msr", false}, {"output-ignored-msrs", false},
    {"comments", true}, {"version", false}, {"help-module", true},
    {"project-file", true}, {"binary-stn-file", true}, {"binary-msr-file", true}, {"seg-file", true}, {"output-database-ids", false},
    {"update-orig-stn-file", false}, {"inversion-method", true}, {"output-json", false},
    {"export-xml-stn-file", false}, {"export-xml-msr-file", false}, {"export-dna-stn-file", false}, {"export-dna-msr-file", false},
    {"gpus", true}, {"first-gpu", true},   // this program's own: shard the adjustment over N GPUs of the node
};

// Boost.program_options accepts unambiguous prefixes (CI uses --phased, --multi)
const Flag* match(const std::string& key, std::string& err)
{
    const Flag* found = nullptr;
    for (const Flag& f : kFlags) {
        if (key == f.name)
            return &f;
        if (std::strncmp(f.name, key.c_str(), key.size()) == 0) {
            if (found) {
                err = "option '--" + key + "' is ambiguous";
                return nullptr;
            }
            found = &f;
        }
    }
    if (!found)
        err = "unrecognised option '--" + key + "'";
    return found;
}

}  // namespace

static std::atomic<bool> g_cancel{false};
extern "C" void sigint_handler(int) { g_cancel.store(true); }

// one option, by its full name (command line and project file share this)
static int apply_option(adjust_settings& s, bool& quiet, const std::string& n, const std::string& value)
{
    if (n == "help" || n == "help-module") {
        std::cout << "usage: dnaadjust <network> [options]      (options may be shortened to an unambiguous prefix)\n\n"
                     "  -n network-name  -i input-folder  -o output-folder  -p project-file  -s binary-stn-file  -m binary-msr-file\n\n";
        int col = 0;
        for (const Flag& f : kFlags) {
            std::string item = std::string("--") + f.name + (f.takes_value ? " arg" : "");
            if (col + (int)item.size() + 2 > 100) {
                std::cout << "\n";
                col = 0;
            }
            std::cout << "  " << item;
            col += (int)item.size() + 2;
        }
        std::cout << "\n\nThe option groups and their meaning are those of DynAdjust's dnaadjust (see INTEGRATION.md).\n";
        return 1;
    } else if (n == "phased-adjustment")
        s.adjust_mode = s.adjust_mode == Phased_Block_1Mode ? Phased_Block_1Mode : PhasedMode;
    else if (n == "multi-thread") {        // both imply a phased adjustment (WRAP:597-621); the execution strategy itself is the GPU's
        s.adjust_mode = s.adjust_mode == Phased_Block_1Mode ? Phased_Block_1Mode : PhasedMode;
        s.multi_thread = !s.stage;
    } else if (n == "staged-adjustment") {
        s.adjust_mode = s.adjust_mode == Phased_Block_1Mode ? Phased_Block_1Mode : PhasedMode;
        s.stage = true;
        s.multi_thread = false;
    }
    else if (n == "gpus") {
        s.gpus = std::atoi(value.c_str());
        if (s.gpus < 1 || s.gpus > 8) {
            std::cerr << "--gpus takes 1 to 8 (the GPUs of one NVLink node)\n";
            return 2;
        }
    } else if (n == "first-gpu")
        s.first_device = std::atoi(value.c_str());
    else if (n == "block1-phased")
        s.adjust_mode = Phased_Block_1Mode;
    else if (n == "simultaneous-adjustment")
        s.adjust_mode = SimultaneousMode;
    else if (n == "conf-interval")
        s.confidence_interval = std::atof(value.c_str());
    else if (n == "iteration-threshold")
        s.iteration_threshold = (double)(float)std::atof(value.c_str());
    else if (n == "max-iterations")
        s.max_iterations = (uint32_t)std::atoi(value.c_str());
    else if (n == "free-stn-sd")
        s.free_std_dev = std::atof(value.c_str());
    else if (n == "fixed-stn-sd")
        s.fixed_std_dev = std::atof(value.c_str());
    else if (n == "scale-normals-to-unity")
        s.scale_normals_to_unity = true;
    else if (n == "input-folder")
        s.input_folder = value;
    else if (n == "output-folder")
        s.output_folder = value;
    else if (n == "output-adj-msr")
        s.output_adj_msr = true;
    else if (n == "output-stn-blocks")
        s.output_stn_blocks = true;
    else if (n == "output-msr-blocks")
        s.output_msr_blocks = true;
    else if (n == "type-b-sd-global")
        s.type_b_global = value;
    else if (n == "type-b-sd-file")
        s.type_b_file = value;
    else if (n == "export-sinex-file")
        s.export_sinex = true;
    else if (n == "output-pos-uncertainty")
        s.output_pos_uncertainty = true;
    else if (n == "output-corrections-file")
        s.output_corrections = true;
    else if (n == "output-apu-vcv-units")
        s.apu_vcv_enu = value == "ENU" || value == "enu" || value == "1";
    else if (n == "hz-corr-threshold")
        s.hz_corr_threshold = std::atof(value.c_str());
    else if (n == "vt-corr-threshold")
        s.vt_corr_threshold = std::atof(value.c_str());
    else if (n == "output-all-covariances")
        s.output_pu_covariances = true;
    else if (n == "report-results")
        s.report_results = true;
    else if (n == "stage-path")
        s.stage_path = value;
    else if (n == "network-name")
        s.network_name = value;
    else if (n == "binary-stn-file")
        s.bst_file = value;
    else if (n == "binary-msr-file")
        s.bms_file = value;
    else if (n == "seg-file")
        s.seg_file = value;
    else if (n == "quiet")
        quiet = true;
    else if (n == "no-binary-update")
        s.update_binary_files = s.update_project_file = false;
    else if (n == "constraints")
        s.station_constraints = value;
    else if (n == "sort-adj-msr-field")
        s.sort_adj_msr = std::atoi(value.c_str());
    else if (n == "output-adj-gnss-units")
        s.adj_gnss_units = std::atoi(value.c_str());
    else if (n == "output-tstat-adj-msr")
        s.adj_msr_tstat = true;
    else if (n == "output-msr-to-stn")
        s.output_msr_to_stn = true;
    else if (n == "sort-msr-to-stn-field")
        s.sort_msr_to_stn = std::atoi(value.c_str());
    else if (n == "stn-corrections")
        s.stn_corrections = true;
    else if (n == "stn-coord-types")
        s.stn_coord_types = value;
    else if (n == "sort-stn-orig-order")
        s.sort_stn_orig_order = true;
    else if (n == "angular-stn-type")
        s.angular_type_stn = std::atoi(value.c_str());
    else if (n == "angular-msr-type")
        s.angular_type_msr = std::atoi(value.c_str());
    else if (n == "dms-msr-format")
        s.dms_format_msr = std::atoi(value.c_str());
    else if (n == "precision-stn-linear")
        s.precision_metres_stn = std::atoi(value.c_str());
    else if (n == "precision-stn-angular")
        s.precision_seconds_stn = std::atoi(value.c_str());
    else if (n == "precision-msr-linear")
        s.precision_metres_msr = std::atoi(value.c_str());
    else if (n == "precision-msr-angular")
        s.precision_seconds_msr = std::atoi(value.c_str());
    else if (n == "output-iter-adj-stn")
        s.iter_adj_stn = true;
    else if (n == "output-iter-adj-stat")
        s.iter_adj_stat = true;
    else if (n == "output-iter-adj-msr")
        s.iter_adj_msr = true;
    else if (n == "output-iter-cmp-msr")
        s.iter_cmp_msr = true;
    else if (n == "output-ignored-msrs")
        s.output_ignored_msrs = true;
    else if (n == "comments")
        s.comments = value;
    else if (n == "output-database-ids")
        s.database_ids = true;
    else if (n == "output-json")
        s.output_json = true;
    else if (n == "export-xml-stn-file")
        s.export_xml_stn = true;
    else if (n == "export-xml-msr-file")
        s.export_xml_msr = true;
    else if (n == "export-dna-stn-file")
        s.export_dna_stn = true;
    else if (n == "export-dna-msr-file")
        s.export_dna_msr = true;
    else if (n == "version") {
        std::cout << "dnaadjust (dynadjust_b200) 1.0\n";
        return 1;
    }
    // remaining accepted flags select CPU execution strategies: no effect here
    return 0;
}

// <net>.dnaproj (CDnaProjectFile::LoadProjectFile dnaprojectfile.cpp:127-310): "variable" in the first 35 columns, its value
// after; sections #general, #adjust and #output hold what dnaadjust reads.  Switches are yes / no.
static int load_project_file(const std::string& file, adjust_settings& s, bool& quiet)
{
    std::ifstream in(file);
    if (!in) {
        std::cerr << "\n- Error: project file " << file << " does not exist.\n\n";
        return 2;
    }
    std::string line, section;
    auto trim = [](std::string t) {
        const size_t a = t.find_first_not_of(" \t\r"), b = t.find_last_not_of(" \t\r");
        return a == std::string::npos ? std::string() : t.substr(a, b - a + 1);
    };
    while (std::getline(in, line)) {
        if (line.size() <= 35 && line.find('#') == std::string::npos)
            continue;
        if (line[0] == '#') {
            section = trim(line.substr(0, line.find(' ')));
            continue;
        }
        if (line.find("----------") != std::string::npos || (section != "#general" && section != "#adjust" && section != "#output"))
            continue;
        std::string var = trim(line.substr(0, 35)), val = line.size() > 35 ? trim(line.substr(35)) : std::string();
        if (val.empty())
            continue;
        if (var == "adjustment-mode") {
            var = val;
            val = "yes";
        }
        const Flag* f = nullptr;
        for (const Flag& k : kFlags)
            if (var == k.name)
                f = &k;
        if (!f || var == "help" || var == "version" || var == "help-module" || var == "project-file")
            continue;   // settings of the other programs of the suite
        if (!f->takes_value) {
            std::string low = val;
            for (char& ch : low)
                ch = (char)std::tolower((unsigned char)ch);
            if (low != "yes" && low != "1" && low != "true")
                continue;
        }
        if (int rc = apply_option(s, quiet, f->name, val))
            return rc;
    }
    return 0;
}

// After an adjustment the project file carries the settings it ran with (CDnaProjectFile::UpdateSettingsAdjust /
// UpdateSettingsOutput / PrintProjectFile, dnaprojectfile.cpp:1703-2027; WRAP:1456-1466): the #adjust and #output sections
// are rewritten, the sections of the other programs of the suite are kept as they are.
static void update_project_file(const adjust_settings& s)
{
    const std::string file = s.output_folder + "/" + s.network_name + ".dnaproj";
    std::vector<std::pair<std::string, std::string>> sections;   // name, text (header line included)
    std::string preamble;
    {
        std::ifstream in(file);
        std::string line;
        while (in && std::getline(in, line)) {
            if (line.size() > 1 && line[0] == '#' && line[1] != ' ')
                sections.emplace_back(line.substr(0, line.find(' ')), std::string());
            if (sections.empty())
                preamble += line + "\n";
            else
                sections.back().second += line + "\n";
        }
    }
    std::ostringstream adj, out;
    const std::string dash(80, '-');
    auto rec = [](std::ostream& os, const std::string& k, const std::string& v) { os << std::left << std::setw(35) << k << std::setw(45) << v << "\n"; };
    auto yn = [](bool b) { return std::string(b ? "yes" : "no"); };
    auto num = [](double v, int prec, bool sci = false) {
        std::ostringstream t;
        t << (sci ? std::scientific : std::fixed) << std::setprecision(prec) << v;
        return t.str();
    };
    auto plain = [](double v) {
        std::ostringstream t;
        t << v;
        return t.str();
    };
    rec(adj, "#adjust (35)", "VALUE");
    adj << dash << "\n";
    rec(adj, "seg-file", s.seg_file.empty() && s.adjust_mode != SimultaneousMode ? s.network_name + ".seg" : s.seg_file);
    rec(adj, "comments", s.comments);
    rec(adj, "adjustment-mode", s.adjust_mode == PhasedMode ? "phased-adjustment" : (s.adjust_mode == Phased_Block_1Mode ? "block1-phased" : "simultaneous-adjustment"));
    rec(adj, "multi-thread", yn(s.multi_thread));
    rec(adj, "staged-adjustment", yn(s.stage));
    rec(adj, "conf-interval", plain(s.confidence_interval));
    rec(adj, "iteration-threshold", plain((float)s.iteration_threshold));
    rec(adj, "max-iterations", std::to_string(s.max_iterations));
    rec(adj, "constraints", s.station_constraints);
    rec(adj, "free-stn-sd", num(s.free_std_dev, 3));
    rec(adj, "fixed-stn-sd", num(s.fixed_std_dev, 4, true));
    rec(adj, "scale-normals-to-unity", yn(s.scale_normals_to_unity));
    rec(adj, "create-stage-files", "no");
    rec(adj, "purge-stage-files", "no");
    rec(adj, "type-b-sd-global", s.type_b_global);
    rec(adj, "type-b-sd-file", s.type_b_file);
    adj << "\n";
    rec(out, "#output (35)", "VALUE");
    out << dash << "\n";
    rec(out, "output-msr-to-stn", yn(s.output_msr_to_stn));
    rec(out, "sort-msr-to-stn-field", std::to_string(s.sort_msr_to_stn));
    rec(out, "output-iter-adj-stn", yn(s.iter_adj_stn));
    rec(out, "output-iter-adj-stat", yn(s.iter_adj_stat));
    rec(out, "output-iter-adj-msr", yn(s.iter_adj_msr));
    rec(out, "output-iter-cmp-msr", yn(s.iter_cmp_msr));
    rec(out, "output-adj-msr", yn(s.output_adj_msr));
    rec(out, "output-adj-gnss-units", std::to_string(s.adj_gnss_units));
    rec(out, "output-tstat-adj-msr", yn(s.adj_msr_tstat));
    rec(out, "sort-adj-msr-field", std::to_string(s.sort_adj_msr));
    rec(out, "output-database-ids", yn(s.database_ids));
    rec(out, "output-msr-blocks", yn(s.output_msr_blocks));
    rec(out, "sort-stn-orig-order", yn(s.sort_stn_orig_order));
    rec(out, "stn-coord-types", s.stn_coord_types);
    rec(out, "angular-stn-type", std::to_string(s.angular_type_stn));
    rec(out, "stn-corrections", yn(s.stn_corrections));
    rec(out, "precision-stn-linear", std::to_string(s.precision_metres_stn));
    rec(out, "precision-stn-angular", std::to_string(s.precision_seconds_stn));
    rec(out, "precision-msr-linear", std::to_string(s.precision_metres_msr));
    rec(out, "precision-msr-angular", std::to_string(s.precision_seconds_msr));
    rec(out, "angular-msr-type", std::to_string(s.angular_type_msr));
    rec(out, "dms-msr-format", std::to_string(s.dms_format_msr));
    rec(out, "output-pos-uncertainty", yn(s.output_pos_uncertainty));
    rec(out, "output-all-covariances", yn(s.output_pu_covariances));
    rec(out, "output-apu-vcv-units", s.apu_vcv_enu ? "1" : "0");
    rec(out, "output-corrections-file", yn(s.output_corrections));
    rec(out, "hz-corr-threshold", num(s.hz_corr_threshold, 3));
    rec(out, "vt-corr-threshold", num(s.vt_corr_threshold, 3));
    rec(out, "export-xml-stn-file", yn(s.export_xml_stn));
    rec(out, "export-dna-stn-file", yn(s.export_dna_stn));
    rec(out, "export-sinex-file", yn(s.export_sinex));
    out << "\n";
    if (sections.empty()) {
        std::ostringstream g;
        preamble = "# " + s.network_name + " project file. Created by dnaadjust (dynadjust_b200).\n\n\n";
        rec(g, "#general (35)", "VALUE");
        g << dash << "\n";
        rec(g, "network-name", s.network_name);
        rec(g, "input-folder", s.input_folder);
        rec(g, "output-folder", s.output_folder);
        rec(g, "verbose-level", "0");
        rec(g, "quiet", "no");
        rec(g, "project-file", file);
        g << "\n";
        sections.emplace_back("#general", g.str());
    }
    bool has_adj = false, has_out = false;
    for (auto& sec : sections) {
        if (sec.first == "#adjust")
            sec.second = adj.str(), has_adj = true;
        if (sec.first == "#output")
            sec.second = out.str(), has_out = true;
    }
    auto before_plot = [&](const std::string& name, const std::string& text) {
        auto it = sections.begin();
        while (it != sections.end() && it->first != "#plot" && it->first != "#display" && !(name == "#adjust" && it->first == "#output"))
            ++it;
        sections.insert(it, {name, text});
    };
    if (!has_adj)
        before_plot("#adjust", adj.str());
    if (!has_out)
        before_plot("#output", out.str());
    std::ofstream os(file);
    os << preamble;
    for (const auto& sec : sections)
        os << sec.second;
}

int main(int argc, char** argv)
{
    adjust_settings s;
    for (int i = 0; i < argc; ++i)
        s.command_line += std::string(argv[i]) + " ";
    bool quiet = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.size() == 2 && a[0] == '-' && a[1] != '-') {   // short forms (WRAP:822-841 and the generic options)
            const char* lng = nullptr;
            switch (a[1]) {
            case 'n': lng = "--network-name"; break;
            case 'i': lng = "--input-folder"; break;
            case 'o': lng = "--output-folder"; break;
            case 'p': lng = "--project-file"; break;
            case 's': lng = "--binary-stn-file"; break;
            case 'm': lng = "--binary-msr-file"; break;
            case 'h': lng = "--help"; break;
            case 'v': lng = "--version"; break;
            }
            if (!lng) {
                std::cerr << "- Error: unrecognised option '" << a << "'\n";
                return EXIT_FAILURE;
            }
            a = lng;
        }
        if (a.rfind("--", 0) != 0) {
            if (!s.network_name.empty()) {
                std::cerr << "- Error: too many positional arguments ('" << a << "').\n";
                return EXIT_FAILURE;
            }
            s.network_name = a;
            continue;
        }
        std::string key = a.substr(2), value;
        size_t eq = key.find('=');
        bool has_eq = eq != std::string::npos;
        if (has_eq) {
            value = key.substr(eq + 1);
            key = key.substr(0, eq);
        }
        std::string err;
        const Flag* f = match(key, err);
        if (!f) {
            std::cerr << "- Error: " << err << "\n";   // WRAP:1040-1046
            return EXIT_FAILURE;
        }
        if (f->takes_value && !has_eq) {
            if (i + 1 >= argc) {
                std::cerr << "- Error: the required argument for option '--" << f->name << "' is missing\n";
                return EXIT_FAILURE;
            }
            value = argv[++i];
        }
        if (std::string(f->name) == "project-file") {
            // "If specified, all other options are ignored" (WRAP:487-505)
            s = adjust_settings();
            s.command_line = std::string(argv[0]) + " -p " + value + " ";
            if (int rc = load_project_file(value, s, quiet))
                return rc == 1 ? EXIT_SUCCESS : EXIT_FAILURE;
            break;
        }
        if (int rc = apply_option(s, quiet, f->name, value))
            return rc == 1 ? EXIT_SUCCESS : EXIT_FAILURE;
    }
    if (s.network_name.empty()) {
        std::cerr << "- Error: no network name was given.\n";
        return EXIT_FAILURE;
    }
    try {
        dna_adjust adj;
        if (s.report_results || s.max_iterations < 1) {   // WRAP:607-614
            if (!quiet)
                std::cout << "+ Report last adjustment results\n";
            adj.LoadLastAdjustment(s);
            adj.PrintAdjustedNetwork();
            return EXIT_SUCCESS;
        }
        std::signal(SIGINT, sigint_handler);   // graceful cancellation between iterations (WRAP:67-71, 1362)
        adj.SetCancelFlag(&g_cancel);
        if (!quiet)
            std::cout << "+ Preparing for adjustment... " << std::flush;
        adj.PrepareAdjustment(s);
        if (!quiet)
            std::cout << "done.\n+ Adjusting network (" << adj.Info().nstations << " stations, " << adj.Info().nfronts
                      << " fronts)..." << std::endl;
        ADJUST_STATUS st = adj.AdjustNetwork();
        if (st == ADJUST_CANCELLED) {
            adj.PrintFailedAdjustment();
            std::cout << "\n- Adjustment cancelled by the user after " << adj.CurrentIteration() << " iteration(s).\n";
            return EXIT_SUCCESS;
        }
        if (st == ADJUST_MAX_ITERATIONS_EXCEEDED) {
            // no statistics or tables for an adjustment that ran out of iterations: the iterations, the status and the
            // stations that kept swinging (WRAP:1386-1390)
            adj.PrintFailedAdjustment();
            if (!quiet)
                std::cout << "+ Solution failed to converge after " << adj.CurrentIteration() << " iteration(s).\n";
            adj.PrintOscillationSummary(std::cout);
            return EXIT_SUCCESS;
        }
        adj.GenerateStatistics();
        adj.SerialiseAdjustedVarianceMatrices();   // <net>-rva.mtx / -pam.mtx for --report-results (WRAP:1397-1399)
        adj.PrintAdjustedNetwork();
        if (s.update_binary_files)
            adj.UpdateBinaryFiles();
        if (!quiet) {
            std::cout << "+ Solution converged after " << adj.CurrentIteration() << " iteration(s).\n";
            std::cout << "+ Chi squared " << adj.GetChiSquared() << ", degrees of freedom " << adj.GetDegreesOfFreedom()
                      << ", rigorous sigma zero " << adj.GetSigmaZero() << "\n";
            adj.PrintOscillationSummary(std::cout);
            adj.PrintSuspectMeasurementSummary(std::cout);   // WRAP:1442-1443
            std::cout << "\n+ Open " << s.network_name << "." << adj.ModeSuffix() << ".adj to view the adjustment details.\n\n";
        }
        if (s.update_project_file)
            update_project_file(s);
        return EXIT_SUCCESS;   // ADJUST_SUCCESS even when not converged: the status is reported in the text (WRAP:133-147)
    } catch (const std::exception& e) {
        std::cerr << "\n- Error: " << e.what() << "\n";
        return EXIT_FAILURE;
    }
}
