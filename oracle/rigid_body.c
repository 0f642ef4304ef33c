/* oracle/rigid_body.c -- C restatement of the free rigid-body step (TEST INFRASTRUCTURE, see oracle/__init__.py).
 *
 * The reference advances the MAV with closed-source PhysX (IsaacGymEnvs/isaacgymenvs/tasks/base/vec_task_asymmetry.py:313);
 * the integrator is OUR specification (oracle/rigid_body.py docstring).  The specification uses
 * fused multiply-adds at fixed places; torch / numpy cannot express a float32 FMA, C can (fmaf is exact by the C
 * standard), so this file is the executable form of the specification and oracle/rigid_body.py:_integrate_py is its
 * slow pure-numpy twin (exact FMA emulation), kept bit-identical by tests/test_oracle_semantics.py.
 *
 * Build (oracle/build_c.py, called by __graft_entry__.build()):
 *   gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC oracle/rigid_body.c -o oracle/librigid_body.so -lm
 * -ffp-contract=off: only the explicit fmaf() calls are fused.
 */
#include <math.h>

typedef struct { float x, y, z; } v3;

static inline v3 cross_f(v3 a, v3 b) {
    v3 r;
    r.x = fmaf(a.y, b.z, -(a.z * b.y));
    r.y = fmaf(a.z, b.x, -(a.x * b.z));
    r.z = fmaf(a.x, b.y, -(a.y * b.x));
    return r;
}

/* One simulate(dt) call for n bodies, in place.  pos/vel/wb/fb/tb: (n,3); quat: (n,4) xyzw.  wb = body-frame angular velocity,
 * fb / tb = body-frame force / torque.  h = dt/substeps, half_h = h/2, half_h2 = (h/2)^2, c_* = -1/6, 1/120, 1/24, all
 * already rounded to float32 by the caller exactly as the CUDA host code rounds them (taco_env.cu: taco_env_create). */
void taco_oracle_integrate(int n, float* pos, float* quat, float* vel, float* wb_, const float* fb_, const float* tb_, float h, float half_h,
                           float half_h2, float c_sin3, float c_sin5, float c_cos4, float inv_mass, int substeps) {
    const float inv_iy = (float)(1.0 / 7e-4);
    for (int i = 0; i < n; ++i) {
        float qx = quat[4 * i], qy = quat[4 * i + 1], qz = quat[4 * i + 2], qw = quat[4 * i + 3];
        v3 p = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
        v3 v = {vel[3 * i], vel[3 * i + 1], vel[3 * i + 2]};
        v3 wb = {wb_[3 * i], wb_[3 * i + 1], wb_[3 * i + 2]};
        const v3 fb = {fb_[3 * i], fb_[3 * i + 1], fb_[3 * i + 2]};
        const v3 tb = {tb_[3 * i], tb_[3 * i + 1], tb_[3 * i + 2]};
        /* F_w = F_b + w t + u x t,  t = 2 (u x F_b): the force is rotated once and held in the world frame */
        const v3 u = {qx, qy, qz};
        v3 t = cross_f(u, fb);
        t.x = t.x + t.x; t.y = t.y + t.y; t.z = t.z + t.z;
        const v3 ut = cross_f(u, t);
        const v3 fw = {fmaf(qw, t.x, fb.x) + ut.x, fmaf(qw, t.y, fb.y) + ut.y, fmaf(qw, t.z, fb.z) + ut.z};
        const v3 acc = {fw.x * inv_mass, fw.y * inv_mass, fmaf(fw.z, inv_mass, -9.81f)};
        for (int s = 0; s < substeps; ++s) {
            v.x = fmaf(h, acc.x, v.x); v.y = fmaf(h, acc.y, v.y); v.z = fmaf(h, acc.z, v.z);
            const v3 iw = {5e-4f * wb.x, 7e-4f * wb.y, 8e-4f * wb.z};
            const v3 gy = cross_f(wb, iw);
            wb.x = fmaf(h, (tb.x - gy.x) * 2000.0f, wb.x);
            wb.y = fmaf(h, (tb.y - gy.y) * inv_iy, wb.y);
            wb.z = fmaf(h, (tb.z - gy.z) * 1250.0f, wb.z);
            p.x = fmaf(h, v.x, p.x); p.y = fmaf(h, v.y, p.y); p.z = fmaf(h, v.z, p.z);
            const float w2 = fmaf(wb.z, wb.z, fmaf(wb.y, wb.y, wb.x * wb.x));
            const float th2 = w2 * half_h2;
            const float kk = half_h * fmaf(th2, fmaf(th2, c_sin5, c_sin3), 1.0f);
            const float cs = fmaf(th2, fmaf(th2, c_cos4, -0.5f), 1.0f);
            const float dx = wb.x * kk, dy = wb.y * kk, dz = wb.z * kk;
            const float rw = fmaf(-qz, dz, fmaf(-qy, dy, fmaf(-qx, dx, qw * cs)));
            const float rx = fmaf(-qz, dy, fmaf(qy, dz, fmaf(qx, cs, qw * dx)));
            const float ry = fmaf(-qx, dz, fmaf(qz, dx, fmaf(qy, cs, qw * dy)));
            const float rz = fmaf(-qy, dx, fmaf(qx, dy, fmaf(qz, cs, qw * dz)));
            const float n2 = fmaf(rw, rw, fmaf(rz, rz, fmaf(ry, ry, rx * rx)));
            const float rn = fmaf(-0.5f, n2, 1.5f);
            qx = rx * rn; qy = ry * rn; qz = rz * rn; qw = rw * rn;
        }
        quat[4 * i] = qx; quat[4 * i + 1] = qy; quat[4 * i + 2] = qz; quat[4 * i + 3] = qw;
        pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
        vel[3 * i] = v.x; vel[3 * i + 1] = v.y; vel[3 * i + 2] = v.z;
        wb_[3 * i] = wb.x; wb_[3 * i + 1] = wb.y; wb_[3 * i + 2] = wb.z;
    }
}
