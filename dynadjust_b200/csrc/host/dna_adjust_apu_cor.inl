// dna_adjust_apu_cor.inl — part of class dna_adjust (included inside the class body by dna_adjust_host.hpp): .apu and .cor reports.

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
        var("Stations printed in blocks:", a_.adjust_mode != SimultaneousMode && (a_.output_pu_covariances || a_.output_stn_blocks) ? "Yes" : "No");
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
            for (const auto& blk : BlockStationLists()) {
                if (blk.first >= 0)
                    os << "Block " << blk.first + 1 << "\n";
                os << header;
                for (uint32_t i : StationOrder(blk.first != -1 ? &blk.second : nullptr))
                    PrintPosUncertainty(os, i);
                if (blk.first >= 0)
                    os << "\n";
            }
            return;
        }
        // --output-all-covariances (PrintPosUncertainty PRN:4438-4484): after each station, its 3x3 covariance blocks with the
        // stations that follow it in the block, from the block's dense variance matrix; phased adjustments list block by block
        const uint32_t nblocks = (uint32_t)info_.nfronts;
        for (uint32_t b = 0; b < nblocks; ++b) {
            if (a_.adjust_mode == Phased_Block_1Mode && b > 0)
                break;
            uint32_t n = 0;
            block_vcv(b, &n, nullptr, 0, nullptr);
            std::vector<uint32_t> st(n);
            const size_t dim = 3 * (size_t)n;
            std::vector<double> q(dim * (dim + 1) / 2);
            block_vcv(b, &n, st.data(), n, q.data());
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
    // the station lists of a report: one list of every station, or — phased modes with --output-stn-blocks — one per .seg
    // block (inner and junction stations); block-1 mode stops after the first (PRN:535-595, 1349-1408, 2665-2770)
    std::vector<std::pair<int, std::vector<uint32_t>>> BlockStationLists() const
    {
        std::vector<std::pair<int, std::vector<uint32_t>>> lists;
        const bool phased = a_.adjust_mode != SimultaneousMode && !seg_.isl.empty();
        if (!phased || (!a_.output_stn_blocks && a_.adjust_mode != Phased_Block_1Mode)) {
            lists.emplace_back(-1, std::vector<uint32_t>());
            return lists;
        }
        for (size_t b = 0; b < seg_.isl.size(); ++b) {
            std::vector<uint32_t> list(seg_.isl[b]);
            if (b < seg_.jsl.size())
                list.insert(list.end(), seg_.jsl[b].begin(), seg_.jsl[b].end());
            std::sort(list.begin(), list.end());
            list.erase(std::unique(list.begin(), list.end()), list.end());
            lists.emplace_back(a_.output_stn_blocks ? (int)b : -2, list);
            if (a_.adjust_mode == Phased_Block_1Mode)
                break;
        }
        return lists;
    }

    void PrintNetworkStationCorrections(const std::string& file) const
    {
        std::ofstream os(file);
        PrintStationFileHeader(os, "CORRECTIONS", file);
        os << std::left << std::setw(35) << "Stations printed in blocks:" << (a_.adjust_mode != SimultaneousMode && a_.output_stn_blocks ? "Yes" : "No")
           << "\n" << std::string(80, '-') << "\n\n";
        os << "Corrections to stations\n------------------------------------------\n\n";
        char buf[512];
        snprintf(buf, sizeof(buf), "%-20s%2s%19s%19s%19s%19s%11s%11s%11s", "Station", "", "Azimuth", "V. Angle", "S. Distance", "H. Distance",
                 "east", "north", "up");
        const std::string header = std::string(buf) + "\n" + std::string(20 + 2 + 4 * 19 + 3 * 11, '-') + "\n";
        for (const auto& blk : BlockStationLists()) {
        if (blk.first >= 0)
            os << "Block " << blk.first + 1 << "\n";
        os << header;
        for (uint32_t i : StationOrder(blk.first != -1 ? &blk.second : nullptr)) {
            const dna_stn_t& s = stn_[i];
            double o[3];
            OriginalXYZ(i, o);
            const double d[3] = {est_[3 * (size_t)i] - o[0], est_[3 * (size_t)i + 1] - o[1], est_[3 * (size_t)i + 2] - o[2]};
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
    }
