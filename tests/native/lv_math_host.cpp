// Test-only host build of the product's device math header (2dliw-slam_b200/csrc/lv_math.cuh): lets the
// CPU test-suite check the closed-form / dual-number Jacobian columns against the oracle without a GPU.
// This is NOT a CPU fallback of the product: it is compiled by tests/test_device_math_host.py only.
#include <cstring>
#include "../../2dliw-slam_b200/csrc/lv_math.cuh"
using namespace lv;
extern "C" {
void lvm_exp(const double* w, double* R) { M3<double> m = exp_so3(load3(w)); std::memcpy(R, m.m, sizeof(m.m)); }
void lvm_log(const double* R, double* w) { M3<double> m; std::memcpy(m.m, R, sizeof(m.m)); V3<double> v = log_so3(m); w[0] = v.x; w[1] = v.y; w[2] = v.z; }
void lvm_imu(const Consts* C, const double* blob, const double* si, const double* sj, double* r_raw, double* J_raw /*[15][30]*/) {
    imu_raw_residual(*C, blob, si, sj, r_raw);
    for (int c = 0; c < 30; ++c) { double col[15]; imu_jacobian_column(*C, blob, si, sj, c, col); for (int r = 0; r < 15; ++r) J_raw[r * 30 + c] = col[r]; }
}
void lvm_wheel(const Consts* C, const double* blob, const double* pi, const double* pj, double* res, double* J /*[3][12]*/) {
    wheel_residuals<double>(*C, blob, load3(pi), load3(pi + 3), load3(pj), load3(pj + 3), res);
    for (int c = 0; c < 12; ++c) {
        V3<Dual> a[4] = {lift<Dual>(load3(pi)), lift<Dual>(load3(pi + 3)), lift<Dual>(load3(pj)), lift<Dual>(load3(pj + 3))};
        V3<Dual>& t = a[c / 3];
        (c % 3 == 0 ? t.x : (c % 3 == 1 ? t.y : t.z)).d = 1.0;
        Dual r[3];
        wheel_residuals<Dual>(*C, blob, a[0], a[1], a[2], a[3], r);
        for (int k = 0; k < 3; ++k) J[k * 12 + c] = r[k].d;
    }
}
void lvm_ground(const Consts* C, const double* pose, double* res, double* J /*[2][6]*/) {
    wheel_residuals<double>;  // (instantiate)
    double rp, rq;
    ground_residuals<double>(*C, load3(pose), load3(pose + 3), &rp, &rq);
    res[0] = rp; res[1] = rq;
    for (int c = 0; c < 6; ++c) {
        V3<Dual> a[2] = {lift<Dual>(load3(pose)), lift<Dual>(load3(pose + 3))};
        V3<Dual>& t = a[c / 3];
        (c % 3 == 0 ? t.x : (c % 3 == 1 ? t.y : t.z)).d = 1.0;
        Dual dp, dq;
        ground_residuals<Dual>(*C, a[0], a[1], &dp, &dq);
        J[c] = dp.d; J[6 + c] = dq.d;
    }
}
// one (frame a, frame b) item through the unified dual evaluation: raw imu residual/Jacobian, wheel, ground
void lvm_item(const Consts* C, const double* imu_blob, const double* wheel_blob, const double* sa, const double* sb, double* r_imu,
              double* J_imu /*[15][30]*/, double* r_wheel, double* J_wheel /*[3][30]*/, double* r_ground, double* J_ground /*[2][30]*/) {
    for (int c = 0; c <= 30; ++c) {
        FrameState<Dual> a = seed_frame_state(sa, c < 15 ? c : -1), b = seed_frame_state(sb, (c >= 15 && c < 30) ? c - 15 : -1);
        Dual ri[15], rw[3], rg[2];
        item_residuals<Dual>(*C, imu_blob, wheel_blob, true, a, b, ri, rw, rg);
        if (c == 30) {
            for (int k = 0; k < 15; ++k) r_imu[k] = ri[k].a;
            for (int k = 0; k < 3; ++k) r_wheel[k] = rw[k].a;
            for (int k = 0; k < 2; ++k) r_ground[k] = rg[k].a;
        } else {
            for (int k = 0; k < 15; ++k) J_imu[k * 30 + c] = ri[k].d;
            for (int k = 0; k < 3; ++k) J_wheel[k * 30 + c] = rw[k].d;
            for (int k = 0; k < 2; ++k) J_ground[k * 30 + c] = rg[k].d;
        }
    }
}
// the closed-form path of factor_kernel: phases 1/2 then one column per "lane"
void lvm_item_analytic(const Consts* C, const double* imu_blob, const double* wheel_blob, const double* sa, const double* sb, double* r_imu,
                       double* J_imu /*[15][30]*/, double* r_wheel, double* J_wheel /*[3][30]*/, double* r_ground, double* J_ground /*[2][30]*/) {
    ItemShared S;
    for (int slot = 0; slot < 3; ++slot) item_phase1(imu_blob, sa, sb, slot, &S);
    item_phase2(*C, imu_blob, sa, sb, 0, &S);
    item_values(*C, imu_blob, wheel_blob, true, S, sa, sb, r_imu, r_wheel, r_ground);
    for (int c = 0; c < 30; ++c) {
        double ci[15], cw[3], cg[2];
        item_column(*C, imu_blob, wheel_blob, true, S, c, ci, cw, cg);
        for (int k = 0; k < 15; ++k) J_imu[k * 30 + c] = ci[k];
        for (int k = 0; k < 3; ++k) J_wheel[k * 30 + c] = cw[k];
        for (int k = 0; k < 2; ++k) J_ground[k * 30 + c] = cg[k];
    }
}
void lvm_frame_table(const Consts* C, const double* pose, double* tab) { laser_frame_table(*C, pose, tab); }
void lvm_so3_plus(const double* t, const double* d, double* o) { so3_plus(t, d, o); }
}
