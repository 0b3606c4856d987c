// dna_adjust_diagnostics.inl — part of class dna_adjust (included inside the class body by dna_adjust_host.hpp): iteration diagnostics: oscillating stations, suspect measurements, failed adjustments.

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
        adj << "\n" << dash << "\n" << std::left << std::setw(35) << "SOLUTION" << (adjustStatus_ == ADJUST_CANCELLED ? "Adjustment cancelled" : "Failed to converge") << "\n";
        char buf[64];
        snprintf(buf, sizeof(buf), "00:00:%09.6f", total_ms_ / 1e3);
        adj << std::left << std::setw(35) << "Total time" << buf << "\n\n";
        std::ofstream xyz(stem + ".xyz");
        PrintOutputFileHeaderInfo(xyz, "DYNADJUST COORDINATE OUTPUT FILE", stem + ".xyz");
    }
