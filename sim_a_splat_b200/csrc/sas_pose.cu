// sas_pose.cu -- the two small per-step pieces either side of the render:
//
// (1) link_pose_kernel: the pose algebra of SplatHandler.draw_handler
//     (/root/reference/sim_a_splat/splat/splat_handler.py:250-281), batched over (env, link) on the
//     device instead of ~10 small NumPy products per link per env in Python:
//         Rm  = R(q / |q|),  tm = p + body_offset  (the weld translation for robot bodies, 0 for free bodies)
//         rot = Ri Rm Rfk^T Ri^T
//         pos = -rot ti - s Ri Rm Rfk^T tfk + s Ri tm + ti
//     evaluated in fp64 (the reference's precision; B200 keeps full-rate-enough FP64 for 8 x 256
//     threads) and rounded once to the fp32 [E,L,12] block the projection kernel reads.
//     Restated on the CPU by oracle/pose_oracle.py:link_transform, which tests/golden pins to the
//     reference's own code.
//
// (2) unproject_kernel: the pinhole unprojection of GaussianSplat.generate_RGBD_point_cloud
//     (/root/reference/sim_a_splat/ns_utils/nerfstudio_utils.py:441-452), the consumer that fixes the
//     depth convention (z along the optical axis): x = (u - cx) d / fx, y = (v - cy) d / fy, z = d,
//     valid = d < max_depth.  HBM bound: 4 B in, 13 B out per pixel.
#include "sas_common.cuh"

namespace {

__global__ void __launch_bounds__(128) link_pose_kernel(const double *__restrict__ body /*[E,L,7] wxyz xyz*/, SasPoseModel m, int E,
                                                        float *__restrict__ out /*[E,L,12]*/) {
    const int idx = blockIdx.x * 128 + threadIdx.x;
    if (idx >= E * m.L) return;
    const int l = idx % m.L;
    const double *b = body + (size_t)idx * 7;
    double w = b[0], x = b[1], y = b[2], z = b[3];
    const double n = sqrt(w * w + x * x + y * y + z * z);
    w /= n; x /= n; y /= n; z /= n;
    const double Rm[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z),     2 * (x * z + w * y),
                          2 * (x * y + w * z),     1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                          2 * (x * z - w * y),     2 * (y * z + w * x),     1 - 2 * (x * x + y * y)};
    const double *Rfk = m.fk_R + 9 * l, *tfk = m.fk_t + 3 * l, *off = m.offset + 3 * l;
    const double tm[3] = {b[4] + off[0], b[5] + off[1], b[6] + off[2]};
    double M1[9], A[9], rot[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) M1[3 * i + j] = m.Ri[3 * i] * Rm[j] + m.Ri[3 * i + 1] * Rm[3 + j] + m.Ri[3 * i + 2] * Rm[6 + j];  // Ri Rm
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[3 * i + j] = M1[3 * i] * Rfk[3 * j] + M1[3 * i + 1] * Rfk[3 * j + 1] + M1[3 * i + 2] * Rfk[3 * j + 2];  // . Rfk^T
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) rot[3 * i + j] = A[3 * i] * m.Ri[3 * j] + A[3 * i + 1] * m.Ri[3 * j + 1] + A[3 * i + 2] * m.Ri[3 * j + 2];  // . Ri^T
    float *o = out + (size_t)idx * 12;
    for (int i = 0; i < 3; i++) {
        const double rti = rot[3 * i] * m.ti[0] + rot[3 * i + 1] * m.ti[1] + rot[3 * i + 2] * m.ti[2];
        const double atf = A[3 * i] * tfk[0] + A[3 * i + 1] * tfk[1] + A[3 * i + 2] * tfk[2];
        const double rtm = m.Ri[3 * i] * tm[0] + m.Ri[3 * i + 1] * tm[1] + m.Ri[3 * i + 2] * tm[2];
        o[9 + i] = (float)(-rti - m.s * atf + m.s * rtm + m.ti[i]);
        for (int j = 0; j < 3; j++) o[3 * i + j] = (float)rot[3 * i + j];
    }
}

// (3) fk_kernel: URDF forward kinematics, batched over envs -- the pose source for users without Drake.
//     The reference obtains body poses from Drake's SceneGraph (manipulator_env.py:377-390) and the capture-time link
//     frames from urchin's visual_trimesh_fk (splat_handler.py:158-163, :197); here joint angles [E,A] go through the
//     URDF's joint tree on the device: child = parent . origin . motion(q), revolute motion by Rodrigues' formula
//     (I + sin q K + (1 - cos q) K^2), prismatic motion axis * q, mimic joints q = m q_src + o -- the same chain
//     sim_a_splat_b200/urdf_fk.py:BatchedBodyPoseSource evaluates with numpy on the host.  One thread per env; the
//     (quaternion wxyz | position) block it writes is what link_pose_kernel reads, so joint angles -> link poses never
//     leave the device.  fp64 throughout.
constexpr int FK_MAX_JOINTS = 48;

__global__ void __launch_bounds__(64) fk_kernel(const double *__restrict__ q /*[E,A]*/, int A, SasFkModel m, int E,
                                                const double *__restrict__ extra /*[E,L-B,7] or null*/, int L,
                                                double *__restrict__ body /*[E,L,7]*/) {
    const int e = blockIdx.x * 64 + threadIdx.x;
    if (e >= E) return;
    double F[FK_MAX_JOINTS + 1][12];   // link frames: rotation row-major (9) | translation (3); frame 0 = root link
    for (int k = 0; k < 12; k++) F[0][k] = m.base[k];
    for (int j = 0; j < m.J; j++) {
        const double *Tp = F[m.parent[j]];
        const double *O = m.origin + 12 * j;
        double qj = 0.0;
        if (m.qindex[j] >= 0) qj = q[(size_t)e * A + m.qindex[j]] * m.mimic[2 * j] + m.mimic[2 * j + 1];
        double Rm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tmv[3] = {0, 0, 0};
        const double ax = m.axis[3 * j], ay = m.axis[3 * j + 1], az = m.axis[3 * j + 2];
        if (m.type[j] == 1) {  // revolute / continuous
            const double sn = sin(qj), c1 = 1.0 - cos(qj);
            const double K[9] = {0, -az, ay, az, 0, -ax, -ay, ax, 0};
            for (int i = 0; i < 3; i++)
                for (int k = 0; k < 3; k++) {
                    const double kk = K[3 * i] * K[k] + K[3 * i + 1] * K[3 + k] + K[3 * i + 2] * K[6 + k];
                    Rm[3 * i + k] = (i == k ? 1.0 : 0.0) + sn * K[3 * i + k] + c1 * kk;
                }
        } else if (m.type[j] == 2) {  // prismatic
            tmv[0] = ax * qj; tmv[1] = ay * qj; tmv[2] = az * qj;
        }
        // child = parent . origin . motion
        double OM[12];
        for (int i = 0; i < 3; i++) {
            for (int k = 0; k < 3; k++) OM[3 * i + k] = O[3 * i] * Rm[k] + O[3 * i + 1] * Rm[3 + k] + O[3 * i + 2] * Rm[6 + k];
            OM[9 + i] = O[3 * i] * tmv[0] + O[3 * i + 1] * tmv[1] + O[3 * i + 2] * tmv[2] + O[9 + i];
        }
        double *Tc = F[j + 1];
        for (int i = 0; i < 3; i++) {
            for (int k = 0; k < 3; k++) Tc[3 * i + k] = Tp[3 * i] * OM[k] + Tp[3 * i + 1] * OM[3 + k] + Tp[3 * i + 2] * OM[6 + k];
            Tc[9 + i] = Tp[3 * i] * OM[9] + Tp[3 * i + 1] * OM[10] + Tp[3 * i + 2] * OM[11] + Tp[9 + i];
        }
    }
    for (int b = 0; b < L; b++) {
        double *o = body + ((size_t)e * L + b) * 7;
        if (b >= m.B) {  // a free body (the T-block): its pose comes from the caller, or stays at the identity
            for (int k = 0; k < 7; k++) o[k] = extra ? extra[((size_t)e * (L - m.B) + (b - m.B)) * 7 + k] : (k == 0 ? 1.0 : 0.0);
            continue;
        }
        const double *T = F[m.body_frame[b]];
        // rotation -> unit quaternion (Shepperd: the largest of w, x, y, z first); link_pose_kernel is sign-agnostic
        const double r00 = T[0], r11 = T[4], r22 = T[8], tr = r00 + r11 + r22;
        double w, x, y, z;
        if (tr > 0.0) {
            const double s4 = 2.0 * sqrt(tr + 1.0);
            w = 0.25 * s4; x = (T[7] - T[5]) / s4; y = (T[2] - T[6]) / s4; z = (T[3] - T[1]) / s4;
        } else if (r00 > r11 && r00 > r22) {
            const double s4 = 2.0 * sqrt(1.0 + r00 - r11 - r22);
            w = (T[7] - T[5]) / s4; x = 0.25 * s4; y = (T[1] + T[3]) / s4; z = (T[2] + T[6]) / s4;
        } else if (r11 > r22) {
            const double s4 = 2.0 * sqrt(1.0 + r11 - r00 - r22);
            w = (T[2] - T[6]) / s4; x = (T[1] + T[3]) / s4; y = 0.25 * s4; z = (T[5] + T[7]) / s4;
        } else {
            const double s4 = 2.0 * sqrt(1.0 + r22 - r00 - r11);
            w = (T[3] - T[1]) / s4; x = (T[2] + T[6]) / s4; y = (T[5] + T[7]) / s4; z = 0.25 * s4;
        }
        o[0] = w; o[1] = x; o[2] = y; o[3] = z; o[4] = T[9]; o[5] = T[10]; o[6] = T[11];
    }
}

__global__ void __launch_bounds__(256) unproject_kernel(const float *__restrict__ depth, int W, int H, float fx, float fy, float cx, float cy,
                                                        float max_depth, float *__restrict__ xyz, unsigned char *__restrict__ valid) {
    const int im = blockIdx.y;
    const size_t P = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < P; i += (size_t)gridDim.x * 256) {
        const int v = (int)(i / W), u = (int)(i - (size_t)v * W);
        const size_t pi = (size_t)im * P + i;
        const float d = depth[pi];
        xyz[3 * pi + 0] = ((float)u - cx) * d / fx;
        xyz[3 * pi + 1] = ((float)v - cy) * d / fy;
        xyz[3 * pi + 2] = d;
        if (valid) valid[pi] = d < max_depth ? 1 : 0;
    }
}

}  // namespace

cudaError_t sas_launch_link_poses(const double *body, const SasPoseModel &m, int E, float *out, cudaStream_t st) {
    const int n = E * m.L;
    if (n <= 0) return cudaSuccess;
    link_pose_kernel<<<(n + 127) / 128, 128, 0, st>>>(body, m, E, out);
    return cudaGetLastError();
}

cudaError_t sas_launch_fk(const double *q, int A, const SasFkModel &m, int E, const double *extra, int L, double *body, cudaStream_t st) {
    if (E <= 0) return cudaSuccess;
    fk_kernel<<<(E + 63) / 64, 64, 0, st>>>(q, A, m, E, extra, L, body);
    return cudaGetLastError();
}
int sas_fk_max_joints() { return FK_MAX_JOINTS; }

cudaError_t sas_launch_unproject(const float *depth, int I, int W, int H, float fx, float fy, float cx, float cy, float max_depth, float *xyz,
                                 unsigned char *valid, cudaStream_t st) {
    const size_t P = (size_t)W * H;
    dim3 grid((unsigned)((P + 255) / 256 < 1184 ? (P + 255) / 256 : 1184), I);
    unproject_kernel<<<grid, 256, 0, st>>>(depth, W, H, fx, fy, cx, cy, max_depth, xyz, valid);
    return cudaGetLastError();
}
