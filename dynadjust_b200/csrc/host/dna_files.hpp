// dna_files.hpp — readers/writers of the DynAdjust files the adjust step consumes and produces (host C++).
//
// Formats follow SURVEY.md Appendix A, restated from the reference's I/O classes:
//   .bst / .bms : 60-byte text header (3 x [10-char label + 10-char value]) + metadata + raw record dump
//                 (include/io/dynadjust_file.cpp:67-116, 190-283; bst_file.cpp:143-178; bms_file.cpp:121-188)
//   .asl        : header + u64 count + {u32 assocMsrCount, u32 amlIndex, u16 validity} per station (asl_file.cpp:80-93)
//   .seg        : ASCII block lists written by dnasegment (seg_file.cpp:57-408, 489-721)
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/dna_records.h"

namespace dnafiles {

constexpr int FIELD = 10;             // identifier_field_width (dynadjust_file.hpp:59)
constexpr int MOD_NAME_WIDTH = 20;    // dnatypes-basic.hpp:75
constexpr int FILE_NAME_WIDTH = 256;  // dnatypes-basic.hpp:76

struct InputFileMeta {
    char filename[FILE_NAME_WIDTH] = {0};
    char epsgCode[DNA_STN_EPSG_WIDTH] = {0};
    char epoch[DNA_STN_EPOCH_WIDTH] = {0};
    char observation_epoch[DNA_STN_EPOCH_WIDTH] = {0};
    uint16_t filetype = 0, datatype = 0;
};

struct BinaryMeta {
    uint64_t binCount = 0;
    bool reduced = false;
    char modifiedBy[MOD_NAME_WIDTH] = {0};
    char epsgCode[DNA_STN_EPSG_WIDTH] = {0};
    char epoch[DNA_STN_EPOCH_WIDTH] = {0};
    char observation_epoch[DNA_STN_EPOCH_WIDTH] = {0};
    bool reftran = false, geoid = false;
    std::vector<InputFileMeta> inputFiles;
    std::vector<std::string> sourceFiles;
    std::string version = "1.2", date, app = "DNA10400";
};

inline std::string trim(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

inline void write_field(std::ofstream& f, const char* label, const std::string& value)
{
    char buf[FIELD + 1];
    f.write(label, FIELD);
    snprintf(buf, sizeof(buf), "%*s", FIELD, value.substr(0, FIELD).c_str());
    f.write(buf, FIELD);
}

inline std::string read_field(std::ifstream& f)
{
    char buf[FIELD + 1];
    buf[FIELD] = 0;
    f.read(buf, FIELD);  // label
    f.read(buf, FIELD);  // value
    return trim(buf);
}

inline bool version_at_least(const std::string& v, int major, int minor)
{
    int a = 0, b = 0;
    sscanf(v.c_str(), "%d.%d", &a, &b);
    return a > major || (a == major && b >= minor);
}

inline void write_header(std::ofstream& f, BinaryMeta& m)
{
    if (m.date.empty()) {
        char d[16];
        time_t t = time(nullptr);
        strftime(d, sizeof(d), "%Y-%m-%d", localtime(&t));
        m.date = d;
    }
    write_field(f, "VERSION   ", m.version);
    write_field(f, "CREATED ON", m.date);
    write_field(f, "CREATED BY", m.app);
}

inline void read_header(std::ifstream& f, BinaryMeta& m)
{
    m.version = read_field(f);
    m.date = read_field(f);
    m.app = read_field(f);
}

inline void write_metadata(std::ofstream& f, const BinaryMeta& m)
{
    f.write(reinterpret_cast<const char*>(&m.binCount), sizeof(uint64_t));
    f.write(reinterpret_cast<const char*>(&m.reduced), sizeof(bool));
    f.write(m.modifiedBy, MOD_NAME_WIDTH);
    f.write(m.epsgCode, DNA_STN_EPSG_WIDTH);
    f.write(m.epoch, DNA_STN_EPOCH_WIDTH);
    f.write(m.observation_epoch, DNA_STN_EPOCH_WIDTH);
    f.write(reinterpret_cast<const char*>(&m.reftran), sizeof(bool));
    f.write(reinterpret_cast<const char*>(&m.geoid), sizeof(bool));
    uint64_t n = m.inputFiles.size();
    f.write(reinterpret_cast<const char*>(&n), sizeof(uint64_t));
    for (const auto& i : m.inputFiles) {
        f.write(i.filename, FILE_NAME_WIDTH);
        f.write(i.epsgCode, DNA_STN_EPSG_WIDTH);
        f.write(i.epoch, DNA_STN_EPOCH_WIDTH);
        f.write(i.observation_epoch, DNA_STN_EPOCH_WIDTH);
        f.write(reinterpret_cast<const char*>(&i.filetype), sizeof(uint16_t));
        f.write(reinterpret_cast<const char*>(&i.datatype), sizeof(uint16_t));
    }
    n = m.sourceFiles.size();
    f.write(reinterpret_cast<const char*>(&n), sizeof(uint64_t));
    for (const auto& s : m.sourceFiles) {
        char name[FILE_NAME_WIDTH] = {0};
        strncpy(name, s.c_str(), FILE_NAME_WIDTH - 1);
        f.write(name, FILE_NAME_WIDTH);
    }
}

inline void read_metadata(std::ifstream& f, BinaryMeta& m)
{
    const bool obs = version_at_least(m.version, 1, 2);
    f.read(reinterpret_cast<char*>(&m.binCount), sizeof(uint64_t));
    f.read(reinterpret_cast<char*>(&m.reduced), sizeof(bool));
    f.read(m.modifiedBy, MOD_NAME_WIDTH);
    f.read(m.epsgCode, DNA_STN_EPSG_WIDTH);
    f.read(m.epoch, DNA_STN_EPOCH_WIDTH);
    if (obs)
        f.read(m.observation_epoch, DNA_STN_EPOCH_WIDTH);
    else
        memcpy(m.observation_epoch, m.epoch, DNA_STN_EPOCH_WIDTH);
    f.read(reinterpret_cast<char*>(&m.reftran), sizeof(bool));
    f.read(reinterpret_cast<char*>(&m.geoid), sizeof(bool));
    uint64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), sizeof(uint64_t));
    if (n > 1000000)
        throw std::runtime_error("corrupt metadata (input file count)");
    m.inputFiles.assign(n, InputFileMeta());
    for (auto& i : m.inputFiles) {
        f.read(i.filename, FILE_NAME_WIDTH);
        f.read(i.epsgCode, DNA_STN_EPSG_WIDTH);
        f.read(i.epoch, DNA_STN_EPOCH_WIDTH);
        if (obs)
            f.read(i.observation_epoch, DNA_STN_EPOCH_WIDTH);
        else
            memcpy(i.observation_epoch, i.epoch, DNA_STN_EPOCH_WIDTH);
        f.read(reinterpret_cast<char*>(&i.filetype), sizeof(uint16_t));
        f.read(reinterpret_cast<char*>(&i.datatype), sizeof(uint16_t));
    }
    m.sourceFiles.clear();
    if (version_at_least(m.version, 1, 1)) {
        f.read(reinterpret_cast<char*>(&n), sizeof(uint64_t));
        if (n > 1000000)
            throw std::runtime_error("corrupt metadata (source file count)");
        for (uint64_t i = 0; i < n; ++i) {
            char name[FILE_NAME_WIDTH + 1] = {0};
            f.read(name, FILE_NAME_WIDTH);
            m.sourceFiles.push_back(name);
        }
    }
}

// LoadFile (bst_file.cpp:100-150 / bms_file.cpp:121-176): header, metadata, then binCount raw records
template <class Rec>
void load_binary(const std::string& path, std::vector<Rec>& recs, BinaryMeta& meta)
{
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f)
        throw std::runtime_error("LoadFile(): An error was encountered when opening " + path + ".");
    read_header(f, meta);
    if (!version_at_least(meta.version, 1, 2))
        throw std::runtime_error("LoadFile(): " + path + " predates observation_epoch support (v1.2); please re-run dnaimport.");
    read_metadata(f, meta);
    const std::string read_error = "LoadFile(): An error was encountered when reading from " + path + ".";
    if (!f)
        throw std::runtime_error(read_error);
    // the record count comes from the file: bound it by what the file can hold before allocating
    const std::streamoff here = f.tellg();
    f.seekg(0, std::ios::end);
    const std::streamoff end = f.tellg();
    f.seekg(here, std::ios::beg);
    if (here < 0 || end < here || meta.binCount > (uint64_t)(end - here) / sizeof(Rec))
        throw std::runtime_error(read_error);
    recs.resize(meta.binCount);
    f.read(reinterpret_cast<char*>(recs.data()), (std::streamsize)(sizeof(Rec) * meta.binCount));
    if (!f)
        throw std::runtime_error(read_error);
}

template <class Rec>
void write_binary(const std::string& path, const std::vector<Rec>& recs, BinaryMeta& meta)
{
    std::ofstream f(path, std::ios::out | std::ios::binary | std::ios::trunc);
    if (!f)
        throw std::runtime_error("WriteFile(): An error was encountered when opening " + path + ".");
    meta.binCount = recs.size();
    write_header(f, meta);
    write_metadata(f, meta);
    f.write(reinterpret_cast<const char*>(recs.data()), (std::streamsize)(sizeof(Rec) * recs.size()));
}

// ---- .asl / .map -----------------------------------------------------------------------------------
// <net>.asl (asl_file.cpp:80-93, dnatemplatestnmsrfuncs.hpp:894-899): header, u64 count, then per station
// { u32 assocMsrCount; u32 amlStnIndex; u16 validity } with no padding.  validity != 0 = the station takes part
// (the reference's parameter list is the stations with validity set, network_data_loader.cpp:146-160).
struct AslEntry {
    uint32_t assoc_msr_count;
    uint32_t aml_index;
    uint16_t validity;
};
inline void load_asl(const std::string& path, std::vector<AslEntry>& asl)
{
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f)
        throw std::runtime_error("LoadFile(): An error was encountered when opening " + path + ".");
    BinaryMeta meta;
    read_header(f, meta);
    uint64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), sizeof(n));
    const std::string read_error = "LoadFile(): An error was encountered when reading from " + path + ".";
    const std::streamoff here = f.tellg();
    f.seekg(0, std::ios::end);
    const std::streamoff end = f.tellg();
    f.seekg(here, std::ios::beg);
    if (!f || here < 0 || n > (uint64_t)(end - here) / 10)
        throw std::runtime_error(read_error);
    asl.resize(n);
    for (AslEntry& e : asl) {
        f.read(reinterpret_cast<char*>(&e.assoc_msr_count), 4);
        f.read(reinterpret_cast<char*>(&e.aml_index), 4);
        f.read(reinterpret_cast<char*>(&e.validity), 2);
    }
    if (!f)
        throw std::runtime_error(read_error);
}

// <net>.map (map_file.cpp:44-98): header, u32 count, then { char name[31]; u32 bstIndex } sorted by name
// (looked up by binary search for --constraints, network_data_loader.cpp:243-244)
inline void load_map(const std::string& path, std::vector<std::pair<std::string, uint32_t>>& map)
{
    std::ifstream f(path, std::ios::in | std::ios::binary);
    if (!f)
        throw std::runtime_error("LoadFile(): An error was encountered when opening " + path + ".");
    BinaryMeta meta;
    read_header(f, meta);
    uint32_t n = 0;
    f.read(reinterpret_cast<char*>(&n), sizeof(n));
    const std::string read_error = "LoadFile(): An error was encountered when reading from " + path + ".";
    const std::streamoff here = f.tellg();
    f.seekg(0, std::ios::end);
    const std::streamoff end = f.tellg();
    f.seekg(here, std::ios::beg);
    if (!f || here < 0 || n > (uint64_t)(end - here) / 35)
        throw std::runtime_error(read_error);
    map.clear();
    map.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        char name[32] = {0};
        uint32_t idx = 0;
        f.read(name, 31);
        f.read(reinterpret_cast<char*>(&idx), 4);
        map.emplace_back(std::string(name, strnlen(name, 31)), idx);
    }
    if (!f)
        throw std::runtime_error(read_error);
}

inline void write_asl(const std::string& path, const std::vector<AslEntry>& asl, BinaryMeta meta)
{
    std::ofstream f(path, std::ios::out | std::ios::binary | std::ios::trunc);
    if (!f)
        throw std::runtime_error("WriteFile(): An error was encountered when opening " + path + ".");
    write_header(f, meta);
    const uint64_t n = asl.size();
    f.write(reinterpret_cast<const char*>(&n), sizeof(n));
    for (const AslEntry& e : asl) {
        f.write(reinterpret_cast<const char*>(&e.assoc_msr_count), 4);
        f.write(reinterpret_cast<const char*>(&e.aml_index), 4);
        f.write(reinterpret_cast<const char*>(&e.validity), 2);
    }
}

inline void write_map(const std::string& path, const std::vector<std::pair<std::string, uint32_t>>& map, BinaryMeta meta)
{
    std::ofstream f(path, std::ios::out | std::ios::binary | std::ios::trunc);
    if (!f)
        throw std::runtime_error("WriteFile(): An error was encountered when opening " + path + ".");
    write_header(f, meta);
    const uint32_t n = (uint32_t)map.size();
    f.write(reinterpret_cast<const char*>(&n), sizeof(n));
    for (const auto& e : map) {
        char name[31] = {0};
        std::memcpy(name, e.first.c_str(), std::min<size_t>(30, e.first.size()));
        f.write(name, 31);
        f.write(reinterpret_cast<const char*>(&e.second), 4);
    }
}

// ---- .seg ------------------------------------------------------------------------------------------
struct Segmentation {
    std::vector<std::vector<uint32_t>> isl, jsl, cml;   // per block: inner stations, junction stations, measurement firsts
    std::vector<uint32_t> net_id;
};

// LoadSegFile (seg_file.cpp:115-408): the block count after "No. blocks produced", one summary row per block
// (fixed-width fields 14,14,16,16,16,16), then per block three 16-wide columns of indices.
inline void load_seg(const std::string& path, Segmentation& seg)
{
    std::ifstream f(path);
    if (!f)
        throw std::runtime_error("LoadSegFile(): An error was encountered when opening " + path + ".");
    const std::string bad = "LoadSegFile(): " + path + " is not a valid segmentation file: ";
    // one unsigned field of a fixed-width row; blank or non-numeric fields are format errors, not exceptions from stoul
    auto field = [&](const std::string& ln, size_t pos, size_t width, const char* what) -> uint32_t {
        const std::string t = pos < ln.size() ? trim(ln.substr(pos, width)) : std::string();
        if (t.empty() || t.size() > 9 || t.find_first_not_of("0123456789") != std::string::npos)
            throw std::runtime_error(bad + what);
        return (uint32_t)std::stoul(t);
    };
    auto next = [&](std::string& ln, const char* what) {
        if (!std::getline(f, ln))
            throw std::runtime_error(bad + "file ends before " + what);
    };
    std::string line;
    uint32_t nblocks = 0;
    while (std::getline(f, line))
        if (line.find("No. blocks produced") != std::string::npos) {
            nblocks = field(line, 35, std::string::npos, "block count");
            break;
        }
    if (!nblocks)
        throw std::runtime_error("LoadSegFile(): no blocks in " + path);
    next(line, "the block summary");  // dashes
    next(line, "the block summary");  // column header
    std::vector<uint32_t> nj(nblocks), ni(nblocks), nm(nblocks);
    seg.net_id.assign(nblocks, 0);
    for (uint32_t b = 0; b < nblocks; ++b) {
        next(line, "the end of the block summary");
        if (field(line, 0, 14, "block number in the summary") != b + 1)
            throw std::runtime_error(bad + "summary rows are not numbered consecutively");
        seg.net_id[b] = field(line, 14, 14, "network id");
        nj[b] = field(line, 28, 16, "junction station count");
        ni[b] = field(line, 44, 16, "inner station count");
        nm[b] = field(line, 60, 16, "measurement count");
        // total must equal inner + junction (seg_file.cpp:277)
        if (field(line, 76, 16, "total station count") != ni[b] + nj[b])
            throw std::runtime_error(bad + "total stations of block " + std::to_string(b + 1) + " is not inner + junction");
    }
    seg.isl.assign(nblocks, {});
    seg.jsl.assign(nblocks, {});
    seg.cml.assign(nblocks, {});
    for (uint32_t b = 0; b < nblocks; ++b) {
        bool found = false;
        while (std::getline(f, line))
            if (line.compare(0, 5, "Block") == 0) {
                found = true;
                break;
            }
        if (!found || field(line, 5, std::string::npos, "block header") != b + 1)
            throw std::runtime_error(bad + "data of block " + std::to_string(b + 1) + " not found");
        next(line, "block data");  // dashes
        for (int k = 0; k < 4; ++k)
            next(line, "block data");  // four count lines
        next(line, "block data");      // blank
        next(line, "block data");      // column header
        next(line, "block data");      // dashes
        uint32_t rows = std::max(std::max(ni[b], nj[b]), nm[b]);
        for (uint32_t r = 0; r < rows; ++r) {
            next(line, "the end of a block's station lists");
            if (r < ni[b])
                seg.isl[b].push_back(field(line, 0, 16, "inner station index"));
            if (r < nj[b])
                seg.jsl[b].push_back(field(line, 16, 16, "junction station index"));
            if (r < nm[b]) {
                // measurement index, optionally followed by its type letter
                std::string d = 32 < line.size() ? trim(line.substr(32, 16)) : std::string();
                size_t e = d.find_first_not_of("0123456789");
                if (e != std::string::npos)
                    d.resize(e);
                if (d.empty() || d.size() > 9)
                    throw std::runtime_error(bad + "measurement index");
                seg.cml[b].push_back((uint32_t)std::stoul(d));
            }
        }
    }
}

inline void write_seg(const std::string& path, const Segmentation& seg, const std::string& bst, const std::string& bms)
{
    std::ofstream f(path);
    const std::string dash(80, '-');
    auto var = [&](const std::string& name, const std::string& value) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%-35s%s", name.c_str(), value.c_str());
        f << buf << "\n";
    };
    f << dash << "\nDYNADJUST SEGMENTATION OUTPUT FILE\n\n";
    var("Version:", "1.2.9 (b200 host)");
    var("Build:", __DATE__);
    var("File created:", "-");
    var("File name:", path);
    f << "\n";
    var("Command line arguments:", "-");
    f << "\n";
    var("Stations file:", bst);
    var("Measurements file:", bms);
    f << "\n";
    var("Minimum inner stations", "150");
    var("Block size threshold", "150");
    var("Starting station(s)", "-");
    f << dash << "\n\nSEGMENTATION SUMMARY\n\n";
    var("No. blocks produced", std::to_string(seg.isl.size()));
    f << dash << "\n";
    char buf[256];
    snprintf(buf, sizeof(buf), "%-14s%-14s%-16s%-16s%-16s%-16s", "Block", "Network ID", "Junction stns", "Inner stns", "Measurements",
             "Total stns");
    f << buf << "\n";
    for (size_t b = 0; b < seg.isl.size(); ++b) {
        snprintf(buf, sizeof(buf), "%-14zu%-14u%-16zu%-16zu%-16zu%-16zu", b + 1, seg.net_id.empty() ? 0u : seg.net_id[b],
                 seg.jsl[b].size(), seg.isl[b].size(), seg.cml[b].size(), seg.isl[b].size() + seg.jsl[b].size());
        f << buf << "\n";
    }
    f << dash << "\n\nINDIVIDUAL BLOCK DATA\n" << dash << "\n";
    for (size_t b = 0; b < seg.isl.size(); ++b) {
        f << "\nBlock " << (b + 1) << "\n" << dash << "\n";
        var("Junction stns:", std::to_string(seg.jsl[b].size()));
        var("Inner stns:", std::to_string(seg.isl[b].size()));
        var("Measurements:", std::to_string(seg.cml[b].size()));
        var("Total stns:", std::to_string(seg.isl[b].size() + seg.jsl[b].size()));
        f << "\n";
        snprintf(buf, sizeof(buf), "%-16s%-16s%-16s", "Inner stns", "Junction stns", "Measurements");
        f << buf << "\n" << dash << "\n";
        size_t rows = std::max(std::max(seg.isl[b].size(), seg.jsl[b].size()), seg.cml[b].size());
        for (size_t r = 0; r < rows; ++r) {
            std::string a = r < seg.isl[b].size() ? std::to_string(seg.isl[b][r]) : "";
            std::string c = r < seg.jsl[b].size() ? std::to_string(seg.jsl[b][r]) : "";
            std::string d = r < seg.cml[b].size() ? std::to_string(seg.cml[b][r]) : "";
            snprintf(buf, sizeof(buf), "%-16s%-16s%-16s", a.c_str(), c.c_str(), d.c_str());
            f << buf << "\n";
        }
        f << dash << "\n";
    }
}

}  // namespace dnafiles
