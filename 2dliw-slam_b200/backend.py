"""Host-side mirror of the reference's back-end optimisation surface on the C ABI.

    keyframe_manager::solve        src/trajectory/keyframe_manager.cpp:722-838
    edge (index1, index2, tf12)    src/trajectory/keyframe_manager.h (seq_edges / loop_edges)
    edge_noise::J                  src/factor/edge_factor.h:4-27

Only the optimisation is mirrored: loop detection (corner / line descriptors, scan matching of loop candidates), the key
frame queue's threading and the map publishing stay with the reference's host code (out of scope).
The compute goes through `Context.pose_graph_solve` (CUDA, csrc/pose_graph.cuh) — there is no CPU path here."""
from dataclasses import dataclass, field

import numpy as np


def edge_noise_J(loop_sigma_p, loop_sigma_q):
    """edge_noise::edge_noise (edge_factor.h:14-26) as written: the second translation weight lands on J(1,2), J(1,1)
    stays 1 (reproduced, not fixed: SURVEY appendix B)."""
    J = np.eye(6)
    J[0, 0] = 1.0 / loop_sigma_p[0]
    J[1, 2] = 1.0 / loop_sigma_p[1]
    J[2, 2] = 1.0 / loop_sigma_p[2]
    J[3, 3], J[4, 4], J[5, 5] = 1.0 / loop_sigma_q[0], 1.0 / loop_sigma_q[1], 1.0 / loop_sigma_q[2]
    return J


@dataclass
class Edge:
    index1: int
    index2: int
    tf12: np.ndarray          # 3x4 (or 4x4) pose of key frame index2 in key frame index1


@dataclass
class KeyFrame:
    p: np.ndarray
    q: np.ndarray


@dataclass
class KeyframeManager:
    """keyframe_queue + seq_edges + loop_edges and `solve()`; poses are written back in place like Ceres does through
    the raw parameter pointers (keyframe_manager.cpp:738-741)."""
    ctx: object                                   # solver.Context created with the back-end's iteration cap (Ceres default 50)
    # defaults: config/corridor.yaml:106 (loop_edge_k), :110-111 (loop_sigma_p / q), :114-115 (use_ground_*_factor)
    loop_sigma_p: tuple = (0.1, 0.1, 0.1)
    loop_sigma_q: tuple = (0.01, 0.01, 0.01)
    loop_edge_k: float = 10.0
    use_ground_p_factor: bool = True
    use_ground_q_factor: bool = True
    keyframe_queue: list = field(default_factory=list)
    seq_edges: list = field(default_factory=list)
    loop_edges: list = field(default_factory=list)
    last_summary: object = None

    def solve(self):
        if not self.keyframe_queue or not (self.seq_edges or self.loop_edges):
            return
        poses = np.array([np.r_[kf.p, kf.q] for kf in self.keyframe_queue], dtype=np.float64)
        edges = self.seq_edges + self.loop_edges          # seq edge 0 first: its index1 is the constant key frame
        index = np.array([(e.index1, e.index2) for e in edges], dtype=np.int32)
        tfs = np.array([np.asarray(e.tf12, dtype=np.float64)[:3, :4] for e in edges])
        weights = np.r_[np.ones(len(self.seq_edges)), np.full(len(self.loop_edges), float(self.loop_edge_k))]
        out, summ = self.ctx.pose_graph_solve(poses, index, tfs, weights, edge_noise_J(self.loop_sigma_p, self.loop_sigma_q),
                                              self.use_ground_p_factor, self.use_ground_q_factor,
                                              fixed_pose=self.seq_edges[0].index1 if self.seq_edges else -1)   # keyframe_manager.cpp:744-748
        for kf, x in zip(self.keyframe_queue, out):
            kf.p[:] = x[0:3]
            kf.q[:] = x[3:6]
        self.last_summary = summ[0]
