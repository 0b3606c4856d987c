// dna_adjust_exports.inl — part of class dna_adjust (included inside the class body by dna_adjust_host.hpp): DNA / DynaML / SINEX exports and the station-table driver.

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
            block_vcv(b, &n, nullptr, 0, nullptr);
            std::vector<uint32_t> st(n);
            const size_t dim = 3 * (size_t)n;
            std::vector<double> q(dim * (dim + 1) / 2);
            block_vcv(b, &n, st.data(), n, q.data());
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
            block_vcv(b, &n, nullptr, 0, nullptr);
            std::vector<uint32_t> st(n);
            const size_t dim = 3 * (size_t)n;
            std::vector<double> q(dim * (dim + 1) / 2);
            block_vcv(b, &n, st.data(), n, q.data());
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
