// gpu_group.hpp — the ranks of a multi-GPU adjustment as threads of this process (dnaadjust --gpus N).
//
// Rank 0 is the context the command line reports from; ranks 1..N-1 are worker threads, each with its own context on its
// own GPU and its own copy of the station / measurement records (the library writes reductions and statistics back into
// the records it borrows).  Every call that contains device-side barriers (prepare + handle exchange, iterate, form the
// inverse, statistics) is made by all ranks at the same time: the main thread posts it to the workers, runs it on rank 0
// and collects the workers' results.  The exchange between the GPUs itself happens on the devices over NVLink
// (include/gadj.h, "Multi-GPU"); this file only keeps the ranks in step — the counterpart of the reference's thread
// pool over blocks (dnaadjust-multi.cpp:92-310).
#pragma once
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/gadj.h"

namespace dynadjust_b200 {

class GpuGroup {
  public:
    ~GpuGroup() { stop(); }

    int size() const { return (int)ranks_.size() + 1; }

    // start the workers: contexts on devices first_device + 1 .. first_device + n - 1, records copied per rank
    void start(int n, int first_device, const gadj_opts& opts, const std::vector<dna_stn_t>& stn, const std::vector<dna_msr_t>& msr)
    {
        for (int r = 1; r < n; ++r) {
            ranks_.emplace_back(new Rank());
            Rank& k = *ranks_.back();
            k.rank = r;
            k.opts = opts;
            k.opts.device = first_device + r;
            k.stn = stn;
            k.msr = msr;
            k.thread = std::thread([this, &k] { loop(k); });
        }
    }

    // run fn(ctx, rank) on every worker rank and main(ctx0) on the calling thread; returns the first error text
    std::string all(const std::function<int(gadj_ctx*, int)>& fn, gadj_ctx* ctx0)
    {
        post(fn);
        std::string err;
        if (fn(ctx0, 0))
            err = gadj_last_error(ctx0);
        std::string werr = wait();
        return err.empty() ? werr : err;
    }
    // run fn on one worker rank only (a getter for something that rank holds)
    std::string one(int rank, const std::function<int(gadj_ctx*, int)>& fn)
    {
        Rank& k = *ranks_[rank - 1];
        {
            std::lock_guard<std::mutex> lk(k.m);
            k.job = fn;
            k.has_job = true;
        }
        k.cv.notify_all();
        std::unique_lock<std::mutex> lk(k.m);
        k.cv.wait(lk, [&] { return !k.has_job; });
        return k.err;
    }
    gadj_ctx* ctx(int rank) { return ranks_[rank - 1]->ctx; }

    void stop()
    {
        for (auto& k : ranks_) {
            {
                std::lock_guard<std::mutex> lk(k->m);
                k->quit = true;
            }
            k->cv.notify_all();
            if (k->thread.joinable())
                k->thread.join();
        }
        ranks_.clear();
    }

  private:
    struct Rank {
        int rank = 0;
        gadj_opts opts{};
        gadj_ctx* ctx = nullptr;
        std::vector<dna_stn_t> stn;
        std::vector<dna_msr_t> msr;
        std::thread thread;
        std::mutex m;
        std::condition_variable cv;
        std::function<int(gadj_ctx*, int)> job;
        bool has_job = false, quit = false;
        std::string err;
    };
    std::vector<std::unique_ptr<Rank>> ranks_;

    void post(const std::function<int(gadj_ctx*, int)>& fn)
    {
        for (auto& k : ranks_) {
            {
                std::lock_guard<std::mutex> lk(k->m);
                k->job = fn;
                k->has_job = true;
            }
            k->cv.notify_all();
        }
    }
    std::string wait()
    {
        std::string err;
        for (auto& k : ranks_) {
            std::unique_lock<std::mutex> lk(k->m);
            k->cv.wait(lk, [&] { return !k->has_job; });
            if (err.empty() && !k->err.empty())
                err = "rank " + std::to_string(k->rank) + ": " + k->err;
        }
        return err;
    }
    void loop(Rank& k)
    {
        if (gadj_create(&k.opts, &k.ctx))
            k.err = gadj_last_error(nullptr);
        for (;;) {
            std::function<int(gadj_ctx*, int)> job;
            {
                std::unique_lock<std::mutex> lk(k.m);
                k.cv.wait(lk, [&] { return k.has_job || k.quit; });
                if (k.quit)
                    break;
                job = k.job;
            }
            std::string e;
            if (!k.ctx)
                e = k.err.empty() ? "no context" : k.err;
            else if (job(k.ctx, k.rank))
                e = gadj_last_error(k.ctx);
            {
                std::lock_guard<std::mutex> lk(k.m);
                k.err = e;
                k.has_job = false;
            }
            k.cv.notify_all();
        }
        if (k.ctx)
            gadj_destroy(k.ctx);
    }

  public:
    // the records of a worker rank (set_stations / set_measurements borrow them)
    dna_stn_t* stn(int rank) { return ranks_[rank - 1]->stn.data(); }
    dna_msr_t* msr(int rank) { return ranks_[rank - 1]->msr.data(); }
};

}  // namespace dynadjust_b200
