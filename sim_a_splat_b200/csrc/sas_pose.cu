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

cudaError_t sas_launch_unproject(const float *depth, int I, int W, int H, float fx, float fy, float cx, float cy, float max_depth, float *xyz,
                                 unsigned char *valid, cudaStream_t st) {
    const size_t P = (size_t)W * H;
    dim3 grid((unsigned)((P + 255) / 256 < 1184 ? (P + 255) / 256 : 1184), I);
    unproject_kernel<<<grid, 256, 0, st>>>(depth, W, H, fx, fy, cx, cy, max_depth, xyz, valid);
    return cudaGetLastError();
}
