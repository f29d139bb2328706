/*
 * dhfk.h -- C ABI of the B200-native DH-AUG hot path (libdhfk.so).
 *
 * The reference (hlz0606/DH-AUG, Python + torch) has no FFI of its own; the path it runs in
 * torch eager is listed below next to the entry point that replaces it.  Citations are
 * relative to DH-AUG_master/ in the reference tree.  INTEGRATION.md shows the ctypes binding
 * a maintainer adds on the reference side.
 *
 *   dhfk_forward / dhfk_backward
 *       Forward_Kinematics_DH_Model.change_3d_joint_angle, torch branch
 *         models_Fk_GAN/forward_kinematics_DH_model.py:562-822   (FK: 33 dh_matrix, 46 bmm)
 *       dh_matrix / rotationMatrix               ...:80-116, :141-191
 *       [:, H36M_32_To_16_Table] gather           models_Fk_GAN/Fk_generator.py:259,453
 *       GAN_torch_world_to_camera                 common/camera.py:36-38 (+ quaternion.py:6-35)
 *       project_to_2d                             common/camera.py:62-94
 *       autograd through all of the above         model_fk_gan_train.py:480 (gen_loss.backward)
 *   dhfk_generator_forward / dhfk_generator_backward          (generator epilogue -> kernel prologue)
 *       the post-MLP glue of Fk_Generator.forward / Video_Fk_Generator.forward,
 *         models_Fk_GAN/Fk_generator.py:121-168 and :310-357 (tanh, tanh*10 root, 31 -> 37 slot scatter,
 *         per-slot range map: ~37 + 37 in-place column writes per call) followed by everything above
 *   dhfk_world_to_camera_*      common/camera.py:36-38 called on its own
 *       (model_fk_gan_train.py:374,434; video_GAN_fun.py:321,440)
 *   dhfk_project_*              common/camera.py:62-94 called on its own, per-row intrinsics
 *       (function_aug/dataloader_update.py:69; model_fk_gan_train.py:376,436)
 *   dhfk_retarget_project       random_bl_aug + project_to_2d of the per-epoch loader refresh
 *       (function_aug/dataloader_update.py:18-41,69; models_Fk_GAN/video_mode_operate.py:879-928)
 *   dhfk_critic_input_* / dhfk_flip_pose   flip, root-centring and KCS features of the critics' inputs
 *       (models_Fk_GAN/Fk_discriminator.py:36-146,269-377; model_fk_gan_train.py:311-331,393-405)
 *   dhfk_video_critic_* / dhfk_video_root_diff_*   per-frame KCS, adjacent-frame differences and playback reverse of
 *       the motion critics' inputs (Fk_discriminator.py:436-512,554-587; video_GAN_fun.py:222-223,269-270)
 *   dhfk_bank_gather            mini-batch out of the device-resident fake-pair bank
 *       (model_fk_gan_train.py:486-510; common/data_loader.py:9-36)
 *   dhfk_grad_allreduce         the gradient exchange of the data-parallel GAN step over NVLink peer memory (SURVEY 8 e)
 *       (the optimizer steps of model_fk_gan_train.py:314-341,382-409,415-482, one replica per GPU)
 *   dhfk_scatter32_*            the [N,32,3] H36M slot layout change_3d_joint_angle returns (...:745-820), both ways
 *   dhfk_topology               the constant tables the reference keeps as Python lists
 *       (forward_kinematics_DH_model.py:234-261,:571-589,:751-817; common/h36m_dataset.py:37-38)
 *
 * Conventions
 *   - All tensors fp32, row-major.  Pointers named *_dev are device pointers owned by the
 *     caller (e.g. the PyTorch caching allocator); the library never allocates, frees or
 *     retains device memory.  `cam` blocks are small HOST arrays read during the call.
 *   - Angles in degrees, lengths / positions in metres.
 *   - Joint order of the 33 angles: right leg 0-4, left leg 5-9, body 10-22, right hand 23-27,
 *     left hand 28-32 (Fk_generator.py:179-184).  Bone order: used_16key_15bone_len_table
 *     (forward_kinematics_DH_model.py:46-49).  Output joints: the 16-joint H36M layout.
 *   - Row strides are in floats and must be >= the logical row length.  Packed, 16-byte
 *     aligned rows take the 128-bit slab path; anything else takes a gather path.
 *     Outputs and upstream gradients ([N,16,3] / [N,16,2]) must be packed and 16-byte aligned.
 *   - Launches are asynchronous on `stream` (a cudaStream_t, e.g.
 *     torch.cuda.current_stream().cuda_stream); no host synchronisation inside.
 *     The caller selects the device (cudaSetDevice / torch.cuda.device).
 *   - Return value: 0 on success, negative DHFK_E_* on argument errors, positive cudaError_t
 *     on CUDA errors.  dhfk_last_error() returns a thread-local message for the last failure.
 *   - Entry points are re-entrant; the library holds no mutable global state (one process-wide, append-only table of
 *     "cudaFuncSetAttribute done" marks aside).
 *   - Non-finite inputs follow the reference's torch semantics: project_to_2d's torch.clamp (common/camera.py:85)
 *     propagates NaN (0/0 at x = z = 0, NaN inputs) and x/0 = +-inf clamps to +-1; the kernels do the same
 *     (FMNMX.NAN).  One documented divergence: 1/z is MUFU.RCP with flush-to-zero, so a DENORMAL depth
 *     (|z| < 1.18e-38) behaves like z = 0 -- (x = 0, z denormal) yields NaN where torch yields 0.
 */
#ifndef DHFK_H_
#define DHFK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DHFK_ABI_VERSION 2

#define DHFK_OK 0
#define DHFK_E_INVAL (-1)       /* null / negative / inconsistent argument                     */
#define DHFK_E_ALIGN (-2)       /* a packed output or gradient pointer is not 16-byte aligned  */
#define DHFK_E_UNSUPPORTED (-3) /* combination not implemented                                 */

/* flags: which sin/cos the fused kernels use (every variant reduces the angle exactly in degrees first).
 *   default (0)              forward: degree-scaled minimax polynomials, max abs err 7.7e-8 (the forward is HBM-bound,
 *                            accuracy is free there); backward: MUFU.SIN/COS, abs err ~4e-7 -- gradients then sit
 *                            1.8e-7 from exact arithmetic (1M poses, relative to max(|ref|,1); the reference's own fp32
 *                            autograd sits 6.7e-8 away, the tolerance is 1e-5) and the backward runs ~10 % faster
 *   DHFK_FLAG_FAST_TRIG      forward also uses MUFU.SIN/COS (positions 1.5e-6 from exact; the reference itself: 1.0e-6)
 *   DHFK_FLAG_ACCURATE_TRIG  backward uses the table sincos (128-entry fp32 table + remainder, max abs err 1.0e-7)
 * The two flags are mutually exclusive.  Numbers: tools/trig_parity.py, tests/test_parity_gpu.py (1M poses). */
#define DHFK_FLAG_FAST_TRIG 0x1u
#define DHFK_FLAG_ACCURATE_TRIG 0x2u

#define DHFK_NUM_JOINTS 33
#define DHFK_NUM_OUT 16
#define DHFK_NUM_BONES 15
#define DHFK_CAM_BLOCK 16 /* host camera block: q(4: w,x,y,z) t(3, metres) f(2) c(2) k(3) p(2)   */

int dhfk_abi_version(void);
const char* dhfk_last_error(void);

/* Constant tables, in joint order 0..32 (any pointer may be NULL).  alpha33 / theta0_33 in degrees. */
int dhfk_topology(int32_t* parent33, int32_t* out16, float* alpha33, float* theta0_33, int32_t* len_kind33,
                  int32_t* len_bone33, int32_t* len_sign33, int32_t* h36m_32_to_16);

/* Poses per CTA of the fused kernels (rows are processed in tiles of this many). */
int dhfk_tile_rows(void);

/*
 * Fused forward:  angles, global rotation, bone lengths, root  ->  world-space 16 joints,
 * optionally camera-space joints and 2D keypoints.
 *   ang_dev   [N, >=33] joint angles (deg), row stride ang_stride
 *   grot_dev  [N, >=3]  global rotation angles x,y,z (deg), row stride grot_stride
 *   bone_dev  [N, >=15] bone lengths (m), row stride bone_stride
 *   root_dev  [N, >=3]  root translation (m), row stride root_stride
 *   cam       host, 16 floats (DHFK_CAM_BLOCK); required iff out_cam_dev or out_uv_dev.  The camera of the fused path is
 *             batch-uniform, as in the GAN step (one subject/camera draw per batch, model_fk_gan_train.py:344-363);
 *             per-row intrinsics (the loader refresh, function_aug/dataloader_update.py:69) go through
 *             dhfk_project_* / dhfk_retarget_project.  (ABI 1 carried a reserved cam_rows_dev pair here; ABI 2 drops it.)
 *   out_world_dev [N,16,3] required; out_cam_dev [N,16,3] / out_uv_dev [N,16,2] optional
 */
int dhfk_forward(const float* ang_dev, int64_t ang_stride, const float* grot_dev, int64_t grot_stride,
                 const float* bone_dev, int64_t bone_stride, const float* root_dev, int64_t root_stride,
                 const float* cam, float* out_world_dev, float* out_cam_dev, float* out_uv_dev, int64_t n,
                 uint32_t flags, void* stream);

/*
 * Fused backward (recomputes the forward in registers):
 *   d/d(ang, grot, root[, bone]) of  <g_world, world> + <g_cam, cam> + <g_uv, uv>.
 *   g_world_dev [N,16,3], g_cam_dev [N,16,3], g_uv_dev [N,16,2]: any may be NULL (= zero), at
 *   least one must be given.  g_ang_dev [N, >=33] (columns 0..32 written), g_grot_dev [N,>=3],
 *   g_root_dev [N,>=3] required; g_bone_dev [N,>=15] optional (the reference never
 *   differentiates bone lengths).
 */
int dhfk_backward(const float* ang_dev, int64_t ang_stride, const float* grot_dev, int64_t grot_stride,
                  const float* bone_dev, int64_t bone_stride, const float* root_dev, int64_t root_stride,
                  const float* cam, const float* g_world_dev, const float* g_cam_dev, const float* g_uv_dev,
                  float* g_ang_dev, int64_t g_ang_stride, float* g_grot_dev, int64_t g_grot_stride,
                  float* g_root_dev, int64_t g_root_stride, float* g_bone_dev, int64_t g_bone_stride,
                  int64_t n, uint32_t flags, void* stream);

/*
 * Generator-epilogue mode.  net_out_dev [N, >=35] is the RAW output of the generator's last Linear layer
 * (before any activation).  Per pose the kernel forms
 *     slot_i  = tanh(net_out[col(i)]) * gen_half37[i] + gen_mid37[i]      for the 31 fed slots i of 37
 *     slot_i  = gen_mid37[i]                                              for slots 4, 9, 22, 23, 28, 33
 *     root    = tanh(net_out[32..34]) * root_scale                         (column 31 is never consumed)
 * with joint angles = slots 0..32 and global rotation = slots 34..36, then runs the same fused
 * FK / camera / projection as dhfk_forward.  gen_half37 / gen_mid37 are HOST arrays: (hi-lo)/2 and (hi+lo)/2
 * of GAN_angle_range_table / GAN_global_rotation_table (Fk_generator.py:35-76) as the generator applies them,
 * or 180 / 0 when GAN_whether_use_preAngle is off; root_scale = 10 (Fk_generator.py:122).
 * bone_dev holds the already scaled lengths boneLength * (1 + scaler) (Fk_generator.py:216-230).
 * The backward writes d/d(net_out) into g_net_out_dev [N, >=35] (column 31 = 0).
 */
int dhfk_generator_forward(const float* net_out_dev, int64_t net_out_stride, const float* bone_dev, int64_t bone_stride,
                           const float* gen_half37, const float* gen_mid37, float root_scale, const float* cam,
                           float* out_world_dev, float* out_cam_dev, float* out_uv_dev, int64_t n, uint32_t flags,
                           void* stream);
int dhfk_generator_backward(const float* net_out_dev, int64_t net_out_stride, const float* bone_dev, int64_t bone_stride,
                            const float* gen_half37, const float* gen_mid37, float root_scale, const float* cam,
                            const float* g_world_dev, const float* g_cam_dev, const float* g_uv_dev,
                            float* g_net_out_dev, int64_t g_net_out_stride, int64_t n, uint32_t flags, void* stream);

/* world -> camera for P = prod(X.shape[:-1]) points with one camera q[4] (w,x,y,z), t[3]:
 * out = qrot(conj(q), X - t).  Backward: g_x = R(q) g_out.
 * cam_on_device = 0: cam_q / cam_t are host arrays read during the call;
 * cam_on_device = 1: they are device pointers read by the kernel (no host sync; this is what the
 * reference call sites have: cam_R / cam_t already live on the GPU, model_fk_gan_train.py:365-366). */
int dhfk_world_to_camera_forward(const float* x_dev, const float* cam_q, const float* cam_t, int32_t cam_on_device,
                                 float* out_dev, int64_t num_points, void* stream);
int dhfk_world_to_camera_backward(const float* g_out_dev, const float* cam_q, int32_t cam_on_device,
                                  float* g_x_dev, int64_t num_points, void* stream);

/* project_to_2d with per-row intrinsics: x_dev [N, J, 3], cam_rows_dev [N, >=9] = f2 c2 k3 p2
 * (row stride cam_rows_stride: 9 or 16 in the reference), uv_dev [N, J, 2]. */
int dhfk_project_forward(const float* x_dev, const float* cam_rows_dev, int64_t cam_rows_stride, float* uv_dev,
                         int64_t n, int64_t joints, void* stream);
int dhfk_project_backward(const float* x_dev, const float* cam_rows_dev, int64_t cam_rows_stride,
                          const float* g_uv_dev, float* g_x_dev, int64_t n, int64_t joints, void* stream);

/*
 * The reference's 32-slot output layout (Forward_Kinematics_DH_Model.change_3d_joint_angle returns [N,32,3],
 * forward_kinematics_DH_model.py:745-820): the 16 joints in their H36M slots (common/h36m_dataset.py:37-38), slot 14 =
 * the head joint again, every other slot = root.  forward: world16_dev [N,16,3] + root_dev [N, root_stride] -> world32_dev
 * [N,32,3].  backward: g_world32_dev -> g_world16_dev [N,16,3] (slot 14 folded into the head joint) and g_root_dev [N,3]
 * = the sum over the 15 free slots (the part of d/d root that does not already flow through world16).
 */
int dhfk_scatter32_forward(const float* world16_dev, const float* root_dev, int64_t root_stride, float* world32_dev,
                           int64_t n, void* stream);
int dhfk_scatter32_backward(const float* g_world32_dev, float* g_world16_dev, float* g_root_dev, int64_t n, void* stream);

/*
 * SURVEY 8 f3 -- per-epoch dataset re-augmentation, fused (forward only; the reference detaches the result):
 *   random_bl_aug            function_aug/dataloader_update.py:18-41
 *   video_mode_random_bl_aug models_Fk_GAN/video_mode_operate.py:879-897
 *     (root-centre; unit bone vectors utils/gan_utils.py:90-134; times a bone-length template row; rebuild the
 *      pose down the 16-joint tree utils/gan_utils.py:56-86; add the root back)
 *   followed by project_to_2d with per-row intrinsics (dataloader_update.py:69, video_mode_operate.py:925-928).
 * pose_dev [N,16,3] packed; templates_dev [T,15] in utils/gan_utils.py bone order (the order of
 * data_extra/bone_length_npy/hm36s15678_bl_templates.npy); tmpl_idx_dev [N] int32 template row per pose, or NULL:
 * every pose uses template row 0 (the video variant draws one row per sequence).  cam_rows_dev [N, >=9] with row
 * stride cam_rows_stride, or stride 0: one intrinsics row shared by all poses.  out_pose_dev [N,16,3] (may alias
 * pose_dev); out_uv_dev [N,16,2] or NULL (then cam_rows_dev is ignored).
 */
int dhfk_retarget_project(const float* pose_dev, const int32_t* tmpl_idx_dev, const float* templates_dev,
                          int32_t num_templates, const float* cam_rows_dev, int64_t cam_rows_stride,
                          float* out_pose_dev, float* out_uv_dev, int64_t n, void* stream);

/*
 * SURVEY 8 f2 -- critic input transforms, fused (one thread per pose):
 *   left/right flip      model_fk_gan_train.py:320-331 (negate x, swap joints [4,5,6,10,11,12] <-> [1,2,3,13,14,15])
 *   root-centring        model_fk_gan_train.py:295,312,437 (x - x[:, :1])
 *   KCS features         Fk_discriminator.py:36-146 special_KCS_Input_transform: 15 bone-pair cosines followed by the
 *                        15 bone lengths (kcs_cols = 30), or Fk_discriminator.py:269-377
 *                        video_mode_special_KCS_Input_transform: the 15 cosines only (kcs_cols = 15); bones from
 *                        special_operate.py:513-539 Fk_get_boneVecByPose3d (used_16key_15bone_len_table order)
 * forward : pos' = centre(flip(pose)) (each step only if its flag is set), kcs = KCS(pos').
 *           out_pos_dev [N,16,3] or NULL; out_kcs_dev [N,kcs_cols] iff kcs_cols > 0.
 * backward: g_pose = d( <g_pos, pos'> + <g_kcs, kcs> ) / d pose; g_pos_dev or g_kcs_dev may be NULL (= zero), not both.
 * jvp     : (t_pos, t_kcs) = J(pose) v_pose -- the derivative of `backward` w.r.t. its upstream gradients, which is
 *           what double-backward through the critic needs (WGAN-GP, Fk_discriminator.py:208-233, create_graph=True).
 * All tensors packed, 16-byte aligned.  out_pos_dev may alias pose_dev.
 */
#define DHFK_CRITIC_CENTRE 0x1u
#define DHFK_CRITIC_FLIP 0x2u
int dhfk_critic_input_forward(const float* pose_dev, float* out_pos_dev, float* out_kcs_dev, int32_t kcs_cols,
                              int64_t n, uint32_t flags, void* stream);
int dhfk_critic_input_backward(const float* pose_dev, const float* g_pos_dev, const float* g_kcs_dev,
                               int32_t kcs_cols, float* g_pose_dev, int64_t n, uint32_t flags, void* stream);
int dhfk_critic_input_jvp(const float* pose_dev, const float* v_pose_dev, float* t_pos_dev, float* t_kcs_dev,
                          int32_t kcs_cols, int64_t n, uint32_t flags, void* stream);
/* The flip alone for [N,16,dims] keypoints, dims = 2 (the 2D critic's inputs, model_fk_gan_train.py:393-405) or 3.
 * out_dev may alias x_dev.  The flip is its own transpose: the backward is the same call on the upstream gradient. */
int dhfk_flip_pose(const float* x_dev, float* out_dev, int64_t n, int32_t dims, void* stream);

/*
 * SURVEY 8 f2, video part -- the inputs of the motion critics, fused (one thread per frame, one pass over the poses):
 *   Video_motion_Fk_3D_Discriminator.forward     models_Fk_GAN/Fk_discriminator.py:436-512
 *       per-frame KCS-15 (video_mode_special_KCS_Input_transform, :269-377)                  kcs  [B,F,15]
 *       adjacent-frame KCS differences (a Python loop of F-1 slice writes + clones, :450-461) dkcs [B,F-1,15]
 *       adjacent-frame 3-D differences (the same loop on the poses, :478-492)                 dpos [B,F-1,48]
 *   Video_motion_Fk_2D_Discriminator.forward     :554-587   root-joint 2-D differences      [B,F-1,2]
 *   temporal playback reverse                    models_Fk_GAN/video_GAN_fun.py:222-223,269-270 (torch.flip(dims=[1]))
 * pose_dev [n_rows,16,3] packed with n_rows = B * frames (clip-major, frame-minor, as the generator emits them).
 * DHFK_VIDEO_REVERSE: every output is what the reference computes from torch.flip(x.view(B,F,-1), dims=[1]); the
 * reversal is index math on the outputs (features of the reversed clip = reversed features, differences negated).
 * forward : out_kcs_dev [B,F,15] and out_dkcs_dev [B,F-1,15] required; out_dpos_dev [B,F-1,48] and out_pos_dev
 *           [B,F,48] (the clip in playback order: the 3-D position branch's input) optional.
 * backward: g_pose_dev [n_rows,16,3] = d( <g_kcs,kcs> + <g_dkcs,dkcs> + <g_dpos,dpos> + <g_pos,pos> ) / d pose; any
 *           upstream gradient may be NULL (= zero), at least one must be given.
 * jvp     : the forward's outputs along the tangent v_pose_dev -- the derivative of `backward` w.r.t. its upstream
 *           gradients (WGAN-GP's create_graph=True pass, Fk_discriminator.py:208-233).
 * frames >= 1 (frames == 1: the difference tensors are empty and their pointers may be NULL).
 */
#define DHFK_VIDEO_REVERSE 0x1u
int dhfk_video_critic_forward(const float* pose_dev, int32_t frames, uint32_t flags, float* out_kcs_dev,
                              float* out_dkcs_dev, float* out_dpos_dev, float* out_pos_dev, int64_t n_rows, void* stream);
int dhfk_video_critic_backward(const float* pose_dev, int32_t frames, uint32_t flags, const float* g_kcs_dev,
                               const float* g_dkcs_dev, const float* g_dpos_dev, const float* g_pos_dev,
                               float* g_pose_dev, int64_t n_rows, void* stream);
int dhfk_video_critic_jvp(const float* pose_dev, const float* v_pose_dev, int32_t frames, uint32_t flags,
                          float* t_kcs_dev, float* t_dkcs_dev, float* t_dpos_dev, float* t_pos_dev, int64_t n_rows,
                          void* stream);
/* 2-D motion critic: uv_dev [n_rows,16,2] -> out_diff_dev [B,F-1,2] = uv[b,f+1,0,:] - uv[b,f,0,:] and, optionally,
 * out_uv_dev [B,F,32] = the clip in playback order (the 2-D position branch's input).  The map is linear: `backward`
 * is its transpose (g_uv_dev [n_rows,16,2] from g_diff_dev / g_uv_playback_dev, either may be NULL, not both) and the
 * forward applied to a tangent is its own JVP. */
int dhfk_video_root_diff_forward(const float* uv_dev, int32_t frames, uint32_t flags, float* out_diff_dev,
                                 float* out_uv_dev, int64_t n_rows, void* stream);
int dhfk_video_root_diff_backward(const float* g_diff_dev, const float* g_uv_playback_dev, int32_t frames,
                                  uint32_t flags, float* g_uv_dev, int64_t n_rows, void* stream);

/*
 * SURVEY 8 f4 -- device-resident fake-pair bank.  The reference copies every iteration's pos_3d_cam / uv / cam to
 * host numpy (model_fk_gan_train.py:486-488) and re-serves them through a CPU DataLoader (:504-510;
 * common/data_loader.py:9-36: PoseDataSet).  Here they stay in HBM, one record of rec_floats floats per pose:
 *   [ pose3d 16x3 | pose2d 16x2 | cam (cam_cols floats, padded to the record end) ]      (default 96 floats = 384 B)
 * and this call serves one (shuffled) mini-batch as three packed tensors -- the wire format of
 * PoseDataSet.__getitem__:  out3d[b] = rec[idx[b]][0:48], out2d[b] = rec[idx[b]][48:80], out_cam[b] = rec[idx[b]][80:80+cam_cols].
 * idx_dev [nb] int64 (what torch.randperm yields).  An index outside [0, bank_rows) produces a NaN row.
 * out_cam_dev may be NULL (cameras not wanted).  rec_floats: multiple of 4, >= 80 + cam_cols rounded up to 4.
 */
int dhfk_bank_gather(const float* bank_dev, int64_t rec_floats, int32_t cam_cols, const int64_t* idx_dev, int64_t nb,
                     int64_t bank_rows, float* out3d_dev, float* out2d_dev, float* out_cam_dev, void* stream);

/*
 * SURVEY 8 e -- the exchange step of the data-parallel GAN iteration: average (scale = 1/world) or sum (scale = 1) a
 * flat fp32 gradient buffer over the `world` GPUs of one node, IN PLACE, in one kernel over NVLink peer memory.
 * Every rank makes the same sequence of calls (same n_floats, max_ctas, cta_threads), one call in flight per flag
 * block.  The call has no per-call state on the host -- the call counter the cross-GPU barriers use lives in the flag
 * block and is advanced by the kernel -- so it can be captured in a CUDA graph and replayed.  Nothing here allocates
 * or maps memory: the caller owns
 *   peer_bufs  [world]  HOST array: address, in THIS process, of every rank's buffer range (index = rank; entry `rank`
 *                       is the local one).  Symmetric allocations mapped into every process
 *                       (torch.distributed._symmetric_memory: handle.buffer_ptrs, or cuMem / cudaIpc mappings).
 *   multicast_buf       address of the same range through the node's NVLS multicast object, or NULL.  With it the
 *                       NVSwitch performs the reduction (multimem.ld_reduce) and the replication (multimem.st);
 *                       without it the kernel loads from / stores to every peer itself, summing in rank order.
 *   peer_flags [world]  HOST array: every rank's flag block, DHFK_AR_FLAG_WORDS uint32 words, zeroed once before the
 *                       first call (on every rank, before any rank's first call) and never touched by the caller again.
 *   status_dev          one local uint32 word, zeroed by the caller; becomes non-zero (the number of the call) if a
 *                       cross-GPU wait outlasted timeout_ms (a rank that never launched): the kernel gives up instead
 *                       of hanging the GPU; the buffer contents are then undefined and the flag blocks must be zeroed
 *                       again on every rank before further use.
 * Every element is summed by exactly one rank and written to all of them: the ranks end bit-identical.  n_floats must
 * be a multiple of 4 and every buffer 16-byte aligned.  max_ctas: 1..DHFK_AR_MAX_CTAS CTAs of cta_threads (a multiple
 * of 32 in 32..512) threads; the kernel is meant to run beside the FK kernels on another stream, and small CTAs fit
 * into the register / thread slots those leave free on an SM.  world <= DHFK_AR_MAX_WORLD.
 */
#define DHFK_AR_MAX_WORLD 16
#define DHFK_AR_MAX_CTAS 64
#define DHFK_AR_FLAG_WORDS (DHFK_AR_MAX_CTAS * 2 * DHFK_AR_MAX_WORLD + DHFK_AR_MAX_CTAS)
int dhfk_grad_allreduce(float* const* peer_bufs, float* multicast_buf, uint32_t* const* peer_flags, uint32_t* status_dev,
                        int32_t rank, int32_t world, int64_t n_floats, float scale, int32_t max_ctas, int32_t cta_threads,
                        int64_t timeout_ms, void* stream);

/*
 * Host-buffer end-to-end entry: forward + backward over N poses whose inputs, upstream gradients
 * and results all live in HOST memory (pinned for full speed).  The rows are cut into chunks that move through a
 * three-stage pipeline -- an upload stream (H2D of the chunk's inputs and upstream gradients), a compute stream
 * (fused forward + backward) and a download stream (D2H of world, uv and the gradients), chained per chunk by
 * events -- over a ring of `num_streams` (1..8) device slots, so both PCIe directions stay busy.  `workspace_dev`
 * is caller-owned device scratch of at least dhfk_host_workspace_bytes(chunk_rows, num_streams) bytes.
 * Synchronous: returns when all results are in host memory.  g_world_host / g_uv_host may be NULL to run forward only.
 */
int64_t dhfk_host_workspace_bytes(int64_t chunk_rows, int32_t num_streams);
int dhfk_forward_backward_host(const float* ang_host, const float* grot_host, const float* bone_host,
                               const float* root_host, const float* cam, const float* g_world_host,
                               const float* g_uv_host, float* out_world_host, float* out_uv_host,
                               float* g_ang_host, float* g_grot_host, float* g_root_host, int64_t n,
                               int64_t chunk_rows, int32_t num_streams, void* workspace_dev,
                               int64_t workspace_bytes, uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* DHFK_H_ */
