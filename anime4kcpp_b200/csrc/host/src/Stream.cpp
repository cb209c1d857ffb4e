// Multi-GPU frame stream: independent frames dealt round-robin to the GPUs, finished frames delivered in order.
//
// Same worker / ordering model as the reference's video filter (video/src/Filter.cpp:33-121): a producer pushes frames into
// bounded channels (back-pressure when the workers fall behind), each worker owns its own CUDA session (stream + scratch,
// like the reference's per-thread state) and upscales whole frames, and the consumer receives results strictly in frame
// order through a min-heap keyed on the frame number (AscendingChannel, util/threads/include/AC/Util/Channel.hpp:19-25).
// Frame n goes to device n mod G; there is no data exchange between GPUs, so no collective and no NCCL.
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <queue>
#include <thread>
#include <vector>

#include "../../../../include/acb200.h"

namespace
{
    struct Job
    {
        long long seq = 0;
        const void* src = nullptr; int w = 0, h = 0, c = 0, src_stride = 0, type = 0;
        double factor = 2.0;
        void* dst = nullptr; int dst_stride = 0;
        // planar video frame (planes > 0): cli/src/Main.cpp:183-206 per frame
        int planes = 0, shift = 0;
        acb200_plane fsrc[3] = {}, fdst[3] = {};
    };
    struct Done
    {
        long long seq; int status;
        bool operator>(const Done& o) const { return seq > o.seq; }
    };
}

struct acb200_stream
{
    const acb200_model* model = nullptr;
    int n_devices = 0;
    std::size_t depth = 2;
    struct Lane     // one per device
    {
        std::mutex m;
        std::condition_variable not_empty, not_full;
        std::deque<Job> q;
    };
    std::vector<Lane> lanes;
    std::vector<std::thread> workers;
    std::vector<acb200_session*> sessions;
    std::atomic<bool> closing{ false };

    std::mutex dm;
    std::condition_variable dcv;
    std::priority_queue<Done, std::vector<Done>, std::greater<Done>> finished;
    long long submitted = 0, delivered = 0;

    void work(int lane_idx, acb200_session* s)
    {
        Lane& lane = lanes[lane_idx];
        for (;;)
        {
            Job job;
            {
                std::unique_lock<std::mutex> lock(lane.m);
                lane.not_empty.wait(lock, [&] { return closing || !lane.q.empty(); });
                if (lane.q.empty()) return;
                job = lane.q.front();
                lane.q.pop_front();
            }
            lane.not_full.notify_one();
            const int rc = job.planes > 0
                ? acb200_process_frame_host(s, model, job.fsrc, job.fdst, job.planes, job.type, job.shift, job.factor)
                : acb200_process_host(s, model, job.src, job.w, job.h, job.c, job.src_stride, job.type, job.factor, job.dst, job.dst_stride);
            {
                std::lock_guard<std::mutex> lock(dm);
                finished.push({ job.seq, rc });
            }
            dcv.notify_all();
        }
    }
};

extern "C"
{
    int acb200_frame_owner(long long seq, int n_devices) { return n_devices > 0 ? static_cast<int>(seq % n_devices) : ACB200_EINVAL; }

    int acb200_stream_create(const acb200_model* model, const int* devices, int n_devices, int workers_per_device, int queue_depth, acb200_stream** out)
    {
        if (!out) return ACB200_EINVAL;
        *out = nullptr;
        if (!model || !devices || n_devices <= 0 || workers_per_device <= 0 || queue_depth <= 0) return ACB200_EINVAL;
        acb200_stream* st = new (std::nothrow) acb200_stream;
        if (!st) return ACB200_ENOMEM;
        st->model = model;
        st->n_devices = n_devices;
        st->depth = static_cast<std::size_t>(queue_depth);
        st->lanes = std::vector<acb200_stream::Lane>(n_devices);
        for (int d = 0; d < n_devices; d++)
            for (int k = 0; k < workers_per_device; k++)
            {
                acb200_session* s = nullptr;
                const int rc = acb200_session_create(devices[d], &s);
                if (rc != ACB200_OK)
                {
                    for (acb200_session* x : st->sessions) acb200_session_destroy(x);
                    delete st;
                    return rc;
                }
                st->sessions.push_back(s);
            }
        int idx = 0;
        for (int d = 0; d < n_devices; d++)
            for (int k = 0; k < workers_per_device; k++, idx++)
                st->workers.emplace_back([st, d, idx] { st->work(d, st->sessions[idx]); });
        *out = st;
        return ACB200_OK;
    }

    static int enqueue(acb200_stream* st, Job& job, long long* seq_out);

    int acb200_stream_submit(acb200_stream* st, const void* src, int w, int h, int c, int src_stride, int elem_type, double factor,
                             void* dst, int dst_stride, long long* seq_out)
    {
        if (!st || !src || !dst) return ACB200_EINVAL;
        Job job;
        job.src = src; job.w = w; job.h = h; job.c = c; job.src_stride = src_stride; job.type = elem_type; job.factor = factor; job.dst = dst; job.dst_stride = dst_stride;
        return enqueue(st, job, seq_out);
    }

    int acb200_stream_submit_frame(acb200_stream* st, const acb200_plane* src, const acb200_plane* dst, int planes, int elem_type, int shift, double factor,
                                   long long* seq_out)
    {
        if (!st || !src || !dst || planes < 1 || planes > 3) return ACB200_EINVAL;
        Job job;
        job.planes = planes; job.shift = shift; job.type = elem_type; job.factor = factor;
        for (int i = 0; i < planes; i++) { job.fsrc[i] = src[i]; job.fdst[i] = dst[i]; }
        return enqueue(st, job, seq_out);
    }

    static int enqueue(acb200_stream* st, Job& job, long long* seq_out)
    {
        {
            std::lock_guard<std::mutex> lock(st->dm);
            job.seq = st->submitted++;
        }
        acb200_stream::Lane& lane = st->lanes[acb200_frame_owner(job.seq, st->n_devices)];
        {
            std::unique_lock<std::mutex> lock(lane.m);
            lane.not_full.wait(lock, [&] { return lane.q.size() < st->depth; });
            lane.q.push_back(job);
        }
        lane.not_empty.notify_one();
        if (seq_out) *seq_out = job.seq;
        return ACB200_OK;
    }

    int acb200_stream_next(acb200_stream* st, long long* seq_out, int* status_out)
    {
        if (!st) return ACB200_EINVAL;
        std::unique_lock<std::mutex> lock(st->dm);
        if (st->delivered >= st->submitted) return ACB200_EINVAL;   // nothing in flight
        st->dcv.wait(lock, [&] { return !st->finished.empty() && st->finished.top().seq == st->delivered; });
        const Done d = st->finished.top();
        st->finished.pop();
        st->delivered++;
        if (seq_out) *seq_out = d.seq;
        if (status_out) *status_out = d.status;
        return ACB200_OK;
    }

    void acb200_stream_destroy(acb200_stream* st)
    {
        if (!st) return;
        for (auto& lane : st->lanes)
        {
            { std::lock_guard<std::mutex> lock(lane.m); st->closing = true; }
            lane.not_empty.notify_all();
        }
        for (auto& t : st->workers) t.join();
        for (acb200_session* s : st->sessions) acb200_session_destroy(s);
        delete st;
    }
}
