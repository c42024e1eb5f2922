// lvio2d_backend_shim.h — reference-side binding of the back-end pose-graph entry point of the C ABI (include/lvio2d.h).
//
//   lvio2d_shim::pose_graph_solve  -> body of keyframe_manager::solve (reference src/trajectory/keyframe_manager.cpp:722-838)
//
// The back-end runs on its own thread in the reference (keyframe_manager.cpp:91) and a context is single-threaded, so
// keyframe_manager owns its OWN lvio2d_ctx (created with max_iters = 0 -> Ceres' default 50; the front-end's context may
// be capped at 10 by fast_mode).  Loop detection, the key-frame queue and the map stay the reference's host code.
//
// NOT compiled in this repository: it needs the reference's own headers (Eigen, keyframe_type.h, params.h), which are
// absent from the build image.  `lvio2d_b200.backend.KeyframeManager` is the same logic in Python
// (tests/test_pose_graph_host.py, tests/test_zz_gpu_pose_graph.py).
#pragma once
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

#include "lvio2d.h"
#include "factor/edge_factor.h"         // edge_noise
#include "trajectory/keyframe_type.h"   // edge
#include "trajectory/trajectory_type.h" // frame_info
#include "utilies/params.h"             // PARAM()

namespace lvio2d_shim
{
    inline void pose_graph_solve(lvio2d_ctx *ctx, std::deque<lvio_2d::frame_info::ptr> &keyframe_queue,
                                 const std::vector<lvio_2d::edge::ptr> &seq_edges, const std::vector<lvio_2d::edge::ptr> &loop_edges)
    {
        const int K = (int)keyframe_queue.size(), E = (int)(seq_edges.size() + loop_edges.size());
        if (K == 0 || E == 0)
            return;
        std::vector<double> poses(6 * K), tf(12 * E), weight(E);
        std::vector<int32_t> index(2 * E);
        for (int i = 0; i < K; i++)
            for (int k = 0; k < 3; k++)
            {
                poses[6 * i + k] = keyframe_queue[i]->p(k);
                poses[6 * i + 3 + k] = keyframe_queue[i]->q(k);
            }
        // seq edges first: the first edge's index1 is the constant key frame (keyframe_manager.cpp:744-748)
        int e = 0;
        auto put = [&](const lvio_2d::edge::ptr &ed, double w)
        {
            index[2 * e] = ed->index1;
            index[2 * e + 1] = ed->index2;
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++)
                    tf[12 * e + 4 * r + c] = ed->tf12.matrix()(r, c);
            weight[e++] = w;
        };
        for (const auto &ed : seq_edges)
            put(ed, 1.0); // edge_factor::Create(tf12, 1), :734
        for (const auto &ed : loop_edges)
            put(ed, PARAM(loop_edge_k)); // :780
        double Jn[36];
        const Eigen::Matrix<double, 6, 6> &J = lvio_2d::edge_noise::get_edge_noise()->J; // as the reference builds it
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++)
                Jn[6 * r + c] = J(r, c);
        lvio2d_summary summary;
        const int rc = lvio2d_pose_graph_solve(ctx, K, poses.data(), E, index.data(), tf.data(), weight.data(), Jn,
                                               PARAM(use_ground_p_factor) ? 1 : 0, PARAM(use_ground_q_factor) ? 1 : 0,
                                               seq_edges.empty() ? -1 : seq_edges[0]->index1, &summary); // SetParameterBlockConstant only inside the seq_edges loop, :744-748
        if (rc != LVIO2D_OK)
            throw std::runtime_error(std::string("lvio2d_pose_graph_solve: ") + lvio2d_strerror(rc) + " (" + lvio2d_last_error(ctx) + ")");
        for (int i = 0; i < K; i++) // Ceres writes through the raw parameter pointers
            for (int k = 0; k < 3; k++)
            {
                keyframe_queue[i]->p(k) = poses[6 * i + k];
                keyframe_queue[i]->q(k) = poses[6 * i + 3 + k];
            }
    }
} // namespace lvio2d_shim
