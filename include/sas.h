/*
 * sas.h -- C ABI of libsas_b200.so, the B200 (sm_100a) splat renderer.
 *
 * This is the drop-in boundary for sim_a_splat's per-step splat render.  The
 * reference has no FFI of its own on this path: it reaches its renderers
 * through Python object protocols (all citations are /root/reference/...):
 *
 *   - gsplat side:  pipeline.model.get_outputs_for_camera(cameras, obb_box)
 *                   sim_a_splat/ns_utils/nerfstudio_utils.py:166-172
 *   - viser side:   server.scene.add_gaussian_splats(...)      sim_a_splat/splat/splat_handler.py:106-119, :134-141
 *                   handle.wxyz / handle.position setters       sim_a_splat/splat/splat_handler.py:283-288
 *                   client.get_render(height,width,wxyz,position)
 *                                                               sim_a_splat/env/splat/splat_env_wrapper.py:148-157,
 *                                                               sim_a_splat/splat/splat_handler.py:339-344
 *
 * Each entry point below names the reference interface it replaces.  Plain
 * pointers and sizes only; no C++ exceptions, torch or Python objects cross it.
 * All functions return SAS_OK (0) or a negative SAS_E* code; sas_last_error()
 * gives the text.  All device work is enqueued on the caller's stream (a
 * cudaStream_t passed as void*; NULL = legacy default stream); nothing
 * synchronises except where stated.  One renderer per device; calls on one
 * renderer are not re-entrant; renderers on different devices are independent.
 * There is NO CPU fallback: every entry point fails with SAS_ECUDA when no
 * CUDA device is usable.
 */
#ifndef SAS_H_
#define SAS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAS_OK 0
#define SAS_EINVAL (-1)   /* bad argument                                             */
#define SAS_ECUDA (-2)    /* CUDA runtime error (text in sas_last_error)               */
#define SAS_ENOMEM (-3)   /* allocation failed                                         */
#define SAS_ECAPACITY (-4)/* intersections / images / envs exceed the configured caps  */
#define SAS_ESTATE (-5)   /* call order (e.g. render before set_cameras)               */

#define SAS_MEM_HOST 0
#define SAS_MEM_DEVICE 1
#define SAS_MEM_HOST_ASYNC 2 /* sas_outputs.mem only: host buffers filled without synchronising, see sas_wait_outputs */

#define SAS_TILE 16

typedef struct sas_renderer sas_renderer; /* opaque */

/*
 * Scene in canonical order: background Gaussians (original index order)
 * followed by the copies of link 0..L-1 (splat_handler.py:104-143).  Values are
 * the ACTIVATED parameters the reference's loader exposes
 * (sim_a_splat/splat/splat_utils.py:33-45: scales=exp(.), opacities=sigmoid(.)),
 * plus the SH coefficients gsplat consumes (features_dc ++ features_rest).
 * The library copies everything; the caller keeps ownership of its buffers.
 */
typedef struct {
    int32_t n_gaussians;
    int32_t sh_degree;      /* 0..3 */
    int32_t n_links;        /* L; link ids are 0..L-1 */
    int32_t mem;            /* SAS_MEM_HOST or SAS_MEM_DEVICE: where the arrays below live */
    const float *means;     /* [N,3] */
    const float *quats;     /* [N,4] wxyz, as stored (normalised by the kernel) */
    const float *scales;    /* [N,3] linear */
    const float *opacities; /* [N]   linear */
    const float *sh;        /* [N,(d+1)^2,3] */
    const int32_t *link_id; /* [N]   -1 = background */
} sas_scene_desc;

typedef struct {
    int32_t max_images;     /* I cap per render call                                  */
    int32_t max_envs;       /* E cap                                                  */
    int32_t max_width;      /* <= 4080 (tile coordinates are packed in 8 bits)        */
    int32_t max_height;
    int64_t max_isects;     /* M cap: merged (image,tile,gaussian) intersections per render call; 0 = auto */
    int64_t max_isects_src; /* cap of the bucket entries actually emitted and sorted (background once per VIEW +
                               links per image; with shared cameras several times smaller than M); 0 = max_isects */
    int32_t radius_rule;    /* SAS_RADIUS_TORCH (0, the pinned default): radii = ceil(3.33 sqrt(cov2d)), gsplat's torch
                               implementation; SAS_RADIUS_CUDA_1_5 (1): gsplat 1.5's CUDA rule, the extend also shrinks with
                               the opacity, min(3.33, sqrt(2 ln(255 o))), and o < 1/255 is culled.  Fewer intersections, the
                               same images (a pixel outside either box has alpha < 1/255) */
    int32_t reserved_;
} sas_config;
#define SAS_RADIUS_TORCH 0
#define SAS_RADIUS_CUDA_1_5 1

/* Output buffers of one render call.  Any pointer may be NULL (not produced). */
typedef struct {
    int32_t mem;            /* SAS_MEM_HOST (pinned or pageable; copied D2H, then the  */
                            /* stream is synchronised), SAS_MEM_DEVICE, or             */
                            /* SAS_MEM_HOST_ASYNC (PINNED host buffers; the D2H copies */
                            /* run on a copy stream of the library out of one of two   */
                            /* alternating device staging sets and the call returns    */
                            /* at once: the copy of render k overlaps render k+1)      */
    int32_t nerfstudio_post;/* 1: apply splatfacto get_outputs post-processing:        */
                            /*    rgb=clamp(rgb+(1-alpha)*background,0,1) and          */
                            /*    depth=where(alpha>0,depth,max(depth))  (the          */
                            /*    GaussianSplat.render contract, nerfstudio_utils.py:123-177) */
    float background[3];    /* composited behind the splats                           */
    float *rgb;             /* [I,H,W,3] f32                                          */
    float *depth;           /* [I,H,W]   f32 expected depth (z along the optical axis,
                               the convention of nerfstudio_utils.py:441-452)          */
    float *alpha;           /* [I,H,W]   f32 accumulation                              */
    uint8_t *rgb_u8;        /* [I,H,W,3] u8 = floor(clamp(rgb,0,1)*255+0.5): what
                               client.get_render returns (splat_env_wrapper.py:148-157) */
    int32_t *n_contrib;     /* [I,H,W]   Gaussians accumulated per pixel (diagnostic)  */
    uint16_t *depth_f16;    /* [I,H,W]   IEEE binary16 bits of the expected depth (round to nearest even): 2 instead
                               of 4 bytes per pixel for observation transfers (uint8 RGB + fp16 depth) */
} sas_outputs;

/* Host outputs travel in as few copies as possible: the library renders the requested outputs of a SAS_MEM_HOST /
 * SAS_MEM_HOST_ASYNC call into ONE device staging block laid out as below, and when the caller's host pointers have
 * the same relative layout (one pinned block cut up with sas_packed_layout) the whole observation goes out in a
 * single cudaMemcpyAsync; otherwise one copy per output.  `want` is a mask of SAS_OUT_* bits; offsets[k] is the byte
 * offset of output k (index = bit number; untouched when not wanted), every offset 256-byte aligned. */
#define SAS_OUT_RGB_U8 1
#define SAS_OUT_DEPTH_F16 2
#define SAS_OUT_DEPTH 4
#define SAS_OUT_ALPHA 8
#define SAS_OUT_RGB 16
#define SAS_OUT_NCONTRIB 32
int sas_packed_layout(int n_images, int width, int height, int want, size_t offsets[6], size_t *total_bytes);

/* Images that share a camera (bit-identical viewmat and K) share a VIEW: the background
 * (link_id < 0) is projected, binned and depth-sorted once per view, the link Gaussians once
 * per image, and the two sorted lists of every (image,tile) are merged.  "src" counts below
 * are the work actually done; n_isects is gsplat's per-image intersection count. */
typedef struct {
    int64_t n_isects;       /* M of the last render: merged (image,tile,gaussian) intersections */
    int64_t n_isects_src;   /* bucket entries emitted and sorted (background per view + links per image) */
    int64_t n_visible_src;  /* (row,gaussian) projections with radii > 0 (rows = views, then images) */
    int64_t max_tile_list;  /* longest merged (image,tile) list                        */
    int64_t n_staged;       /* Gaussian records the compositor staged into shared memory
                               (sum over tiles, after early termination)               */
    int32_t overflow;       /* != 0 if an intersection count exceeded max_isects (frame is incomplete) */
    int32_t n_images, n_views, n_gaussians, n_background, width, height, tile_w, tile_h;
    int32_t pad_;
    int64_t cap_isects;     /* current capacities (sas_config.max_isects / max_isects_src, grown by sas_reserve_isects*) */
    int64_t cap_isects_src;
} sas_stats;

/* Replaces: GSplatLoader -> SplatHandler._load_saved_splats + _add_robot_segmented_splats +
 * _setup_scene_splat (splat_handler.py:85-143), i.e. "upload the scene to the renderer". */
int sas_create(const sas_scene_desc *scene, const sas_config *cfg, int device, sas_renderer **out);
void sas_destroy(sas_renderer *r);

/* Replaces: the per-link `handle.wxyz = ...; handle.position = ...` writes of
 * SplatHandler.draw_handler (splat_handler.py:283-288), batched over envs.
 * poses [n_envs, L, 12] fp32 = (rot row-major 3x3 | pos), meaning
 * mu' = rot*mu + pos, Sigma' = rot*Sigma*rot^T for every Gaussian of link l. */
int sas_set_link_poses(sas_renderer *r, const float *poses, int n_envs, int mem, void *stream);
/* Input lifetime (this call and sas_set_link_poses_from_bodies): the array is READ ASYNCHRONOUSLY by a copy enqueued
 * on `stream`.  Host memory must stay valid and unchanged until that copy has run (pageable memory is staged by the
 * driver before the call returns; pinned memory is not); device memory must not be overwritten by work on OTHER
 * streams until an event recorded on `stream` after this call has completed -- the Python renderers take a private
 * copy on the caller's stream first (SplatRenderer.set_link_poses), so their callers may reuse the tensor at once. */

/* Replaces: the pose algebra of SplatHandler.draw_handler itself (splat_handler.py:250-281), run on the
 * device for all envs at once.  The model holds the constants of _load_saved_masks (:62-83: scale_factor,
 * Ri, ti), the weld translation (:229-235; its rotation must be identity, as the reference asserts) and the
 * capture-time visual-mesh FK of _add_robot_meshes (:158-163, :197).  HOST pointers; copied. */
typedef struct {
    int32_t n_links;          /* must equal the scene's L; link l is driven by the l-th robot body of the draw message */
    double scale;             /* s  */
    double Ri[9];             /* row-major */
    double ti[3];
    const double *fk_R;       /* [L,9] row-major; identity for a free body registered at the origin (the T-block) */
    const double *fk_t;       /* [L,3] */
    const double *body_offset;/* [L,3] added to the body position: the weld translation for robot bodies (:250-257),
                                 zero for free bodies (:296-310); NULL = all zero */
} sas_pose_model;
int sas_set_pose_model(sas_renderer *r, const sas_pose_model *model);

/* body poses [n_envs, L, 7] fp64 = (quaternion wxyz as sent -- normalised here, :239-240 -- | position xyz), the
 * fields of lcmt_viewer_draw for the robot's bodies (manipulator_env.py:377-390).  Computes
 * rot = Ri Rm Rfk^T Ri^T, pos = -rot ti - s Ri Rm Rfk^T tfk + s Ri (p + body_offset) + ti in fp64 and stores the
 * fp32 [n_envs, L, 12] block sas_set_link_poses would have been given. */
int sas_set_link_poses_from_bodies(sas_renderer *r, const double *body_wxyz_xyz, int n_envs, int mem, void *stream);

/* Pose source without Drake.  Replaces: ManipulatorSimEnv._generate_draw_msg's body poses
 * (manipulator_env.py:377-390, Drake SceneGraph) for the robot bodies, computed from joint angles by forward kinematics
 * over the URDF's joint tree ON THE DEVICE -- the chain urchin evaluates for the reference's capture-time frames
 * (splat_handler.py:158-163, :197).  HOST pointers; copied.  Joints in parent-before-child order; link frame 0 is the
 * root link (posed at `base`), frame j+1 the child link of joint j. */
typedef struct {
    int32_t n_joints;           /* J <= 48 */
    int32_t n_actuated;         /* A: columns of the joint-angle array */
    int32_t n_bodies;           /* B <= L: link groups 0..B-1 are driven by these bodies, the rest are free bodies */
    const int32_t *parent;      /* [J] frame index of the joint's parent link */
    const int32_t *type;        /* [J] 0 fixed, 1 revolute / continuous, 2 prismatic */
    const int32_t *qindex;      /* [J] column of q (a mimic joint: the column of the joint it follows), -1 = fixed at 0 */
    const double *mimic;        /* [J,2] q_joint = multiplier * q[qindex] + offset; (1, 0) for ordinary joints */
    const double *origin;       /* [J,16] row-major 4x4 <origin> of the joint */
    const double *axis;         /* [J,3] */
    const double *base;         /* [16] row-major 4x4 world pose of the root link; NULL = identity */
    const int32_t *body_frame;  /* [B] frame index of body b */
} sas_fk_model;
int sas_set_fk_model(sas_renderer *r, const sas_fk_model *model);

/* joint angles [n_envs, A] fp64 -> body poses -> link poses (needs sas_set_pose_model as well).  free_bodies
 * [n_envs, L - B, 7] fp64 (quaternion wxyz | position) are the poses of the link groups that are not robot bodies (the
 * T-block); NULL = identity.  Both arrays live in `mem`; the input lifetime rule of sas_set_link_poses applies. */
int sas_set_link_poses_from_joints(sas_renderer *r, const double *q, int n_envs, const double *free_bodies, int mem, void *stream);

/* Replaces: the pinhole unprojection of GaussianSplat.generate_RGBD_point_cloud (nerfstudio_utils.py:441-452):
 * xyz[i,v,u] = ((u-cx) d/fx, (v-cy) d/fy, d) in the camera frame (OpenCV axes), valid = d < max_depth.
 * K row-major [9]; depth [I,H,W], xyz [I,H,W,3], valid [I,H,W] (may be NULL); all in `mem`. */
int sas_unproject_depth(sas_renderer *r, const float *depth, int n_images, int width, int height, const float *K, float max_depth,
                        float *xyz, uint8_t *valid, int mem, void *stream);

/* Replaces: the (height,width,wxyz,position) arguments of client.get_render
 * (splat_env_wrapper.py:148-157) / the Cameras(...) built in GaussianSplat.render
 * (nerfstudio_utils.py:125-138), batched.  HOST pointers.  viewmats [I,16] are
 * world->camera, row-major, OpenCV axes (+z forward); Ks [I,9] row-major;
 * env_of_image [I] selects the env whose link poses image i uses (NULL = all 0). */
int sas_set_cameras(sas_renderer *r, const float *viewmats, const float *Ks, const int32_t *env_of_image,
                    int n_images, int width, int height, void *stream);

/* Replaces: client.get_render(...) per camera (splat_env_wrapper.py:147-158) and
 * model.get_outputs_for_camera (nerfstudio_utils.py:166-172): renders every image
 * set by sas_set_cameras with the current link poses. */
int sas_render(sas_renderer *r, const sas_outputs *out, void *stream);

/* Blocks until the host buffers of a SAS_MEM_HOST_ASYNC render are complete: age 0 = the latest such render, 1 = the
 * one before it (older ones have completed by construction: a staging set is only reused after its copy drained).
 * The reference's counterpart is the blocking return of client.get_render (splat_env_wrapper.py:148-157); the split
 * into issue + wait lets a vectorised env overlap the observation transfer of step k with the render of step k+1.
 * At most two async renders may be outstanding, and their host buffers must differ. */
int sas_wait_outputs(sas_renderer *r, int age);

/* Grow the intersection buffers (pairs, sorted ids) to hold max_isects entries.  A render whose
 * intersection count exceeded the capacity sets sas_stats.overflow; reserve more and re-render. */
int sas_reserve_isects(sas_renderer *r, long long max_isects);
/* The same with separate capacities: max_src bucket entries (16 B each: pairs + their sort buffer), max_all merged
 * ids (4 B each).  Either may be 0 = leave as it is. */
int sas_reserve_isects2(sas_renderer *r, long long max_src, long long max_all);

/* Overflow without a device synchronisation per frame.  The intersection counts stay on the device; every render
 * leaves a snapshot of its counters in a small device ring.  Renders with host outputs bring their snapshot back
 * with the observation: a SAS_MEM_HOST render, and sas_wait_outputs for a SAS_MEM_HOST_ASYNC one, return
 * SAS_ECAPACITY when that frame was incomplete.  For device outputs, sas_poll_overflow reads the ring (it
 * synchronises `stream`): returns 1 and the capacities that would have sufficed if any of the last (up to 4)
 * renders since the previous poll / reserve overflowed, else 0; negative = error. */
int sas_poll_overflow(sas_renderer *r, long long *need_src, long long *need_all, void *stream);

/* Synchronises the device. */
int sas_get_stats(sas_renderer *r, sas_stats *out);

/* Intermediates of the last render, for parity tests (synchronises). */
#define SAS_DBG_RECORDS 1   /* [C*n_bg + I*(N-n_bg), 12] f32: per view the background records, then per image the link
                               records: m2x,m2y,0.5*conic_a,conic_b, 0.5*conic_c,opacity,extent_x,extent_y, r,g,b,depth
                               (extent = conservative screen-space reach of alpha >= 1/255, used only for culling) */
#define SAS_DBG_BINS 2      /* same rows, [.,2] u32: (x0 | y0<<8 | x1<<16 | y1<<24), bits(depth); tile rect, exclusive max */
#define SAS_DBG_TILE_OFFSETS 3 /* [I*T+1] i32 exclusive scan of the merged per-(image,tile) counts = gsplat isect_offsets */
#define SAS_DBG_SORTED_GIDS 4  /* [M] i32 canonical Gaussian index per intersection, (image,tile,depth,gid) order */
#define SAS_DBG_VIEW_OF_IMAGE 5 /* [I] i32 */
#define SAS_DBG_RADII 6     /* same rows as the records, [.,2] i32: radius_x, radius_y (gsplat radii; undefined where culled) */
#define SAS_DBG_LINK_POSES 7 /* [n_envs, L, 12] f32: the link poses in effect (as set, or as composed from body poses) */
int sas_debug_fetch(sas_renderer *r, int what, void *dst_host, size_t bytes);

/* Per-stage device timing (CUDA events on the caller's stream around each stage of sas_render).
 * Stages: 0 project, 1 scan, 2 emit, 3 bucket sort, 4 composite, 5 nerfstudio post.
 * sas_get_stage_ms synchronises and returns the last render's stage durations in ms. */
#define SAS_N_STAGES 6
int sas_set_stage_timing(sas_renderer *r, int enable);
int sas_get_stage_ms(sas_renderer *r, float *ms_out /*[SAS_N_STAGES]*/);

/* Number of kernels the last sas_render enqueued (for bench.py's gpu_launches). */
int sas_last_launch_count(const sas_renderer *r);

const char *sas_last_error(const sas_renderer *r); /* r may be NULL: error of the last failed sas_create */
const char *sas_version(void);

#ifdef __cplusplus
}
#endif
#endif /* SAS_H_ */
