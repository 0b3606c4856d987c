// dna_adjust_json.inl — part of class dna_adjust (included inside the class body by dna_adjust_host.hpp): JSONL siblings of the text reports (--output-json).

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
            for (uint32_t i : StationOrder(nullptr)) {
                const dna_stn_t& s = stn_[i];
                double o[3], R[9];
                OriginalXYZ(i, o);
                local_rotation(s.currentLatitude, s.currentLongitude, R);
                const double d[3] = {est_[3 * (size_t)i] - o[0], est_[3 * (size_t)i + 1] - o[1], est_[3 * (size_t)i + 2] - o[2]};
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
