// dna_adjust_reportmode.inl — part of class dna_adjust (included inside the class body by dna_adjust_host.hpp): precisions of adjusted baselines, <net>-rva.mtx / -pam.mtx, --report-results.

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
        ComputeStationValidity();
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
        if (a_.adj_msr_tstat) {   // the last run may not have asked for them: t = n-stat / sqrt(sigma zero)
            const double sz = std::sqrt(stats_.sigma_zero);
            for (dna_msr_t& m : msr_)
                if (!m.ignore && m.measStart <= 2)
                    m.TStat = std::fabs(sz) < 1.0e-10 ? 0.0 : m.NStat / sz;
        }
        info_.nstations = (uint32_t)stn_.size();
        info_.nfronts = 0;
    }
    bool ReportMode() const { return report_mode_; }
