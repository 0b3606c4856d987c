// dna_adjust_tables.inl — part of class dna_adjust (included inside the class body by dna_adjust_host.hpp): the .adj / .xyz tables: adjusted, computed and ignored measurements, stations, measurements to station.

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

    void ComputeStationValidity()
    {
        valid_.assign(stn_.size(), 0);
        for (size_t i = 0; i < msr_.size();) {
            const size_t span = MeasurementSpan(i);
            if (!msr_[i].ignore)
                for (size_t j = i; j < i + span && j < msr_.size(); ++j) {
                    const dna_msr_t& r = msr_[j];
                    if (r.ignore || (std::strchr("GXY", r.measType) && r.measStart != 0))
                        continue;
                    auto mark = [&](uint32_t sidx) {
                        if (sidx < valid_.size())
                            valid_[sidx] = 1;
                    };
                    mark(r.station1);
                    if (r.measType != 'Y' && !std::strchr("HRIJPQ", r.measType))
                        mark(r.station2);
                    if (r.measType == 'A')
                        mark(r.station3);
                }
            i += span;
        }
    }

    // <net>.asl (asl_file.cpp:80-93): the station validity flags dnaimport derived (LDR:146-160 builds the parameter list
    // from them).  When the file is there and covers the station file its flags are used for the reports; a station it
    // calls valid that no measurement taking part touches stays out (the engine holds it by its a-priori weight only).
    void LoadAssociatedStationList(const std::string& path)
    {
        if (!std::filesystem::exists(path))
            return;
        std::vector<dnafiles::AslEntry> asl;
        dnafiles::load_asl(path, asl);
        if (asl.size() != stn_.size())
            return;   // written for another station file
        size_t differ = 0;
        for (size_t i = 0; i < asl.size(); ++i) {
            const uint8_t v = asl[i].validity != 0;
            differ += v != valid_[i];
            valid_[i] = v && valid_[i];
        }
        asl_differs_ = differ;
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
        case 7:
            if (a_.adj_gnss_units != 0 && !pam_rec_.empty()) {
                // baselines are listed by the n-statistics of the frame they are printed in (UpdateGNSSNstatsForAlternateUnits PRN:4526-4715)
                std::vector<std::pair<double, uint32_t>> k;
                for (uint32_t f : list) {
                    const dna_msr_t& m = msr_[f];
                    double v = 0.0;
                    if (m.measType == 'G' || m.measType == 'X') {
                        size_t j = f;
                        for (uint32_t b = 0; b < std::max<uint32_t>(1u, m.vectorCount1) && j + 2 < msr_.size(); ++b) {
                            MsrRow rows[3];
                            AlternateUnitRows(j, rows);
                            for (int q = 0; q < 3; ++q)
                                if (std::isfinite(rows[q].nstat))
                                    v = std::max(v, std::fabs(rows[q].nstat));
                            j += 3 + 3 * (size_t)msr_[j].vectorCount2;
                        }
                    } else
                        v = LargestOf(f, [](const dna_msr_t& x) { return x.NStat; });
                    k.emplace_back(v, f);
                }
                std::stable_sort(k.begin(), k.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
                for (size_t i = 0; i < k.size(); ++i)
                    list[i] = k[i].second;
                break;
            }
            by_largest([](const dna_msr_t& m) { return m.NStat; });
            break;
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
        MsrRow rows[3];
        AlternateUnitRows(i, rows);
        for (int q = 0; q < 3; ++q)
            PrintMsrRow(os, rows[q], 0);
    }
    void AlternateUnitRows(size_t i, MsrRow* rows) const
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
            rows[q] = row;
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
            // every station a measurement that takes part touches; the others were not adjusted and are not reported
            // (the reference's station lists hold valid stations only, LDR:286-300)
            for (size_t i = 0; i < stn_.size(); ++i)
                if (valid_.empty() || valid_[i])
                    list.push_back((uint32_t)i);
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
