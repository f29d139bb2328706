"""Constant tables of the DH-AUG hot path (data, not code).

Every table is pinned bit-exactly against values extracted from the unmodified reference
(tests/golden/tables.npz, produced by oracle/make_golden.py) in tests/test_tables.py.
Citations are relative to /root/reference/DH-AUG_master.

Joint numbering: the 33 local DH joints in generator order (Fk_generator.py:179-184):
right leg 0-4, left leg 5-9, body 10-22, right hand 23-27, left hand 28-32.
"""
from __future__ import annotations

import numpy as np

NUM_JOINTS = 33
NUM_OUT = 16
NUM_BONES = 15

# models_Fk_GAN/forward_kinematics_DH_model.py:234-261
ALPHA_DEG = np.array(
    [0, -90, -90, 0, 0] + [0, 90, 90, 0, 0] + [0] + [-90] * 11 + [90] + [-90, -90, -90, 0, 0] + [-90, 90, 90, 0, 0],
    dtype=np.float32)
THETA0_DEG = np.array(
    [0, -90, 180, 0, 0] + [180, -90, 0, 0, 0] + [90] + [-90] * 10 + [0, 0] + [-180, -90, 180, 0, 0] + [0, -90, 0, 0, 0],
    dtype=np.float32)
# chain roots have parent -1; both arm chains continue from body[8] = joint 18 (:633,:648)
PARENT = np.array(
    [-1, 0, 1, 2, 3] + [-1, 5, 6, 7, 8] + [-1] + list(range(10, 22)) + [18, 23, 24, 25, 26] + [18, 28, 29, 30, 31],
    dtype=np.int32)
# bone-length slots (:571-589): kind 0 none / 1 DH 'a' / 2 DH 'd'; bone index into the 15-vector in
# used_16key_15bone_len_table order (:46-49) = kwarg order of change_3d_joint_angle (:357-361)
LEN_KIND = np.zeros(33, np.int32)
LEN_BONE = -np.ones(33, np.int32)
LEN_SIGN = np.zeros(33, np.int32)
for _j, _kind, _bone, _sign in (
        (0, 1, 5, 1), (3, 1, 3, 1), (4, 1, 1, 1),          # right leg: hip, thigh, shin
        (5, 1, 4, -1), (8, 1, 2, 1), (9, 1, 0, 1),         # left leg
        (13, 2, 6, 1), (16, 2, 7, 1), (22, 1, 14, 1),      # waist, thorax (d), neck (a)
        (23, 1, 9, -1), (26, 1, 11, 1), (27, 1, 13, 1),    # right shoulder, upper arm, forearm
        (28, 1, 8, 1), (31, 1, 10, 1), (32, 1, 12, 1)):    # left
    LEN_KIND[_j], LEN_BONE[_j], LEN_SIGN[_j] = _kind, _bone, _sign

# joint whose frame origin is output k (:751-817 composed with common/h36m_dataset.py:37-38):
# Hip, RHip,RKnee,RFoot, LHip,LKnee,LFoot, Spine,Thorax,Head, LShoulder,LElbow,LWrist, RShoulder,RElbow,RWrist
OUT16_JOINT = np.array([10, 0, 3, 4, 5, 8, 9, 13, 16, 22, 28, 31, 32, 23, 26, 27], dtype=np.int32)
# common/h36m_dataset.py:37-38
H36M_32_To_16_Table = [0, 1, 2, 3, 6, 7, 8, 12, 13, 15, 17, 18, 19, 25, 26, 27]
# slot 14 ('Neck/Nose') duplicates the head joint (:787-793); it is the only written slot not gathered
H36M_EXTRA_SLOT_14_OUT = 9

# models_Fk_GAN/forward_kinematics_DH_model.py:46-49
used_16key_15bone_len_table = [(5, 6), (2, 3), (4, 5), (1, 2), (0, 4), (0, 1), (0, 7), (7, 8), (8, 10), (8, 13),
                               (10, 11), (13, 14), (11, 12), (14, 15), (8, 9)]
BONE_NAMES = ("left_small_leg_len", "right_small_leg_len", "left_big_leg_len", "right_big_leg_len",
              "left_hip_len", "right_hip_len", "waist_len", "thorax_len", "left_shoulder_len",
              "right_shoulder_len", "left_big_arm_len", "right_big_arm_len", "left_small_arm_len",
              "right_small_arm_len", "neck_len")

# models_Fk_GAN/Fk_generator.py:41-76: GAN_angle_range_table rows joint1..joint34 (lo, hi), degrees
GAN_ANGLE_RANGE = np.array(
    [(-110, 65), (-110, 65), (-110, 180), (-180, 0), (0, 0), (-65, 110), (-65, 110), (-110, 180), (-180, 0), (0, 0)]
    + [(-180, 180)] * 12 + [(0, 0), (0, 0)]
    + [(-155, 65), (-155, 65), (-100, 180), (0, 180), (0, 0), (-65, 155), (-65, 155), (-100, 180), (0, 180), (0, 0)],
    dtype=np.float32)
# Fk_generator.py:35-39
GAN_GLOBAL_ROT_RANGE = np.array([(-180, 180)] * 3, dtype=np.float32)
# generator slots forced to zero (Fk_generator.py:136): network output never reaches them
GAN_ZERO_SLOTS = (4, 9, 22, 23, 28, 33)

# common/h36m_dataset.py:46-87 -- (id, center, focal_length, radial, tangential, res_w, res_h)
H36M_INTRINSICS = (
    ("54138969", (512.54150390625, 515.4514770507812), (1145.0494384765625, 1143.7811279296875),
     (-0.20709891617298126, 0.24777518212795258, -0.0030751503072679043),
     (-0.0009756988729350269, -0.00142447161488235), 1000, 1002),
    ("55011271", (508.8486328125, 508.0649108886719), (1149.6756591796875, 1147.5916748046875),
     (-0.1942136287689209, 0.2404085397720337, 0.006819975562393665),
     (-0.0016190266469493508, -0.0027408944442868233), 1000, 1000),
    ("58860488", (519.8158569335938, 501.40264892578125), (1149.1407470703125, 1148.7989501953125),
     (-0.2083381861448288, 0.25548800826072693, -0.0024604974314570427),
     (0.0014843869721516967, -0.0007599993259645998), 1000, 1000),
    ("60457274", (514.9682006835938, 501.88201904296875), (1145.5113525390625, 1144.77392578125),
     (-0.198384091258049, 0.21832367777824402, -0.008947807364165783),
     (-0.0005872055771760643, -0.0018133620033040643), 1000, 1002),
)
# common/h36m_dataset.py:89-234 -- subject -> 4 x (orientation quaternion wxyz, translation mm)
H36M_EXTRINSICS = {
    "S1": (((0.1407056450843811, -0.1500701755285263, -0.755240797996521, 0.6223280429840088), (1841.1070556640625, 4955.28466796875, 1563.4454345703125)),
           ((0.6157187819480896, -0.764836311340332, -0.14833825826644897, 0.11794740706682205), (1761.278564453125, -5078.0068359375, 1606.2650146484375)),
           ((0.14651472866535187, -0.14647851884365082, 0.7653023600578308, -0.6094175577163696), (-1846.7777099609375, 5215.04638671875, 1491.972412109375)),
           ((0.5834008455276489, -0.7853162288665771, 0.14548823237419128, -0.14749594032764435), (-1794.7896728515625, -3722.698974609375, 1574.8927001953125))),
    "S5": (((0.1467377245426178, -0.162370964884758, -0.7551892995834351, 0.6178938746452332), (2097.3916015625, 4880.94482421875, 1605.732421875)),
           ((0.6159758567810059, -0.7626792192459106, -0.15728192031383514, 0.1189815029501915), (2031.7008056640625, -5167.93310546875, 1612.923095703125)),
           ((0.14291371405124664, -0.12907841801643372, 0.7678384780883789, -0.6110143065452576), (-1620.5948486328125, 5171.65869140625, 1496.43701171875)),
           ((0.5920479893684387, -0.7814217805862427, 0.1274748593568802, -0.15036417543888092), (-1637.1737060546875, -3867.3173828125, 1547.033203125))),
    "S6": (((0.1337897777557373, -0.15692396461963654, -0.7571090459823608, 0.6198879480361938), (1935.4517822265625, 4950.24560546875, 1618.0838623046875)),
           ((0.6147197484970093, -0.7628812789916992, -0.16174767911434174, 0.11819244921207428), (1969.803955078125, -5128.73876953125, 1632.77880859375)),
           ((0.1529948115348816, -0.13529130816459656, 0.7646096348762512, -0.6112781167030334), (-1769.596435546875, 5185.361328125, 1476.993408203125)),
           ((0.5916101336479187, -0.7804774045944214, 0.12832270562648773, -0.1561593860387802), (-1721.668701171875, -3884.13134765625, 1540.4879150390625))),
    "S7": (((0.1435241848230362, -0.1631336808204651, -0.7548328638076782, 0.6188824772834778), (1974.512939453125, 4926.3544921875, 1597.8326416015625)),
           ((0.6141672730445862, -0.7638262510299683, -0.1596645563840866, 0.1177929937839508), (1937.0584716796875, -5119.7900390625, 1631.5665283203125)),
           ((0.14550060033798218, -0.12874816358089447, 0.7660516500473022, -0.6127139329910278), (-1741.8111572265625, 5208.24951171875, 1464.8245849609375)),
           ((0.5912848114967346, -0.7821764349937439, 0.12445473670959473, -0.15196487307548523), (-1734.7105712890625, -3832.42138671875, 1548.5830078125))),
    "S8": (((0.14110587537288666, -0.15589867532253265, -0.7561917304992676, 0.619644045829773), (2150.65185546875, 4896.1611328125, 1611.9046630859375)),
           ((0.6169601678848267, -0.7647668123245239, -0.14846350252628326, 0.11158157885074615), (2219.965576171875, -5148.453125, 1613.0440673828125)),
           ((0.1471444070339203, -0.13377119600772858, 0.7670128345489502, -0.6100369691848755), (-1571.2215576171875, 5137.0185546875, 1498.1761474609375)),
           ((0.5927824378013611, -0.7825870513916016, 0.12147816270589828, -0.14631995558738708), (-1476.913330078125, -3896.7412109375, 1547.97216796875))),
    "S9": (((0.15540587902069092, -0.15548215806484222, -0.7532095313072205, 0.6199594736099243), (2044.45849609375, 4935.1171875, 1481.2275390625)),
           ((0.618784487247467, -0.7634735107421875, -0.14132238924503326, 0.11933968216180801), (1990.959716796875, -5123.810546875, 1568.8048095703125)),
           ((0.13357827067375183, -0.1367100477218628, 0.7689454555511475, -0.6100738644599915), (-1670.9921875, 5211.98583984375, 1528.387939453125)),
           ((0.5879399180412292, -0.7823407053947449, 0.1427614390850067, -0.14794869720935822), (-1696.04345703125, -3827.099853515625, 1591.4127197265625))),
    "S11": (((0.15232472121715546, -0.15442320704460144, -0.7547563314437866, 0.6191070079803467), (2098.440185546875, 4926.5546875, 1500.278564453125)),
            ((0.6189449429512024, -0.7600917220115662, -0.15300633013248444, 0.1255258321762085), (2083.182373046875, -4912.1728515625, 1561.07861328125)),
            ((0.14943228662014008, -0.15650227665901184, 0.7681233882904053, -0.6026304364204407), (-1609.8153076171875, 5177.3359375, 1537.896728515625)),
            ((0.5894251465797424, -0.7818877100944519, 0.13991211354732513, -0.14715361595153809), (-1590.738037109375, -3854.1689453125, 1578.017578125))),
}
TRAIN_SUBJECTS = ("S1", "S5", "S6", "S7", "S8")  # common/h36m_dataset.py:41

# data_extra/bone_length_npy/hm36s15678_bl_templates.npy: (5,15) float32, one row per train subject,
# in utils/gan_utils.py:90-110 get_BoneVecbypose3d bone order.
BONE_TEMPLATES_GANUTILS_ORDER = np.array([
    [0.13294846, 0.44289464, 0.45420563, 0.13294877, 0.44289464, 0.45420682, 0.23347287, 0.25707757,
     0.1818823, 0.15103406, 0.2788833, 0.25173312, 0.15103163, 0.27889305, 0.2517285],
    [0.11931339, 0.42828554, 0.4424443, 0.1193131, 0.42828605, 0.44244403, 0.22431792, 0.2540553,
     0.16618064, 0.14309622, 0.2645848, 0.24862026, 0.14309767, 0.26458368, 0.24862061],
    [0.14261368, 0.48657086, 0.46149334, 0.14261152, 0.48656374, 0.46149373, 0.2622111, 0.2600094,
     0.19247183, 0.14937478, 0.3010085, 0.25791466, 0.14937413, 0.30100808, 0.25791493],
    [0.13587914, 0.4486152, 0.43800974, 0.13587871, 0.4486154, 0.43800804, 0.22624578, 0.2554076,
     0.16872624, 0.13971755, 0.2755636, 0.24729918, 0.13971443, 0.2755719, 0.24729763],
    [0.14653736, 0.45214748, 0.438633, 0.14653677, 0.45214623, 0.43863365, 0.26121622, 0.25102398,
     0.21279596, 0.16931628, 0.2899084, 0.24417701, 0.16931622, 0.2899071, 0.24417761],
], dtype=np.float32)
# gan_utils bone (i,j) list, to permute a template row into used_16key_15bone_len_table order
_GANUTILS_BONES = [(0, 1), (1, 2), (2, 3), (0, 4), (4, 5), (5, 6), (0, 7), (7, 8), (8, 9), (8, 10), (10, 11),
                   (11, 12), (8, 13), (13, 14), (14, 15)]
TEMPLATE_PERM = np.array([_GANUTILS_BONES.index(b) for b in used_16key_15bone_len_table], dtype=np.int64)
BONE_TEMPLATES = BONE_TEMPLATES_GANUTILS_ORDER[:, TEMPLATE_PERM]
# symmetric scaler groups, Fk_generator.py:216-230 (bone index -> scaler column; -1 = unscaled thorax)
BONE_SCALER_GROUP = np.array([0, 0, 1, 1, 2, 2, 3, -1, 4, 4, 5, 5, 6, 6, 7], dtype=np.int64)


def camera_block(subject: str = "S1", cam_id: int = 0) -> np.ndarray:
    """16-float batch-uniform camera block [q(4) wxyz, t(3) metres, f(2), c(2), k(3), p(2)], built the
    way the GAN loop builds cam_R / cam_t / cam_para_temp each iteration
    (models_Fk_GAN/model_fk_gan_train.py:344-363): float64 numpy arithmetic, `c` rounded to float32 by
    normalize_screen_coordinates(...).astype('float32') (common/camera.py:10-15), then float32."""
    (q, t) = H36M_EXTRINSICS[subject][cam_id]
    (_id, center, focal, k, p, res_w, res_h) = H36M_INTRINSICS[cam_id]
    res_w = float(res_w)
    res_h = float(res_h)
    q = np.array(q, dtype=np.float64)
    t = np.array(t, dtype=np.float64) / 1000.0
    f = np.array(focal, dtype=np.float64) / res_w * 2.0
    c = np.array(center, dtype=np.float64)
    c = np.array([c[0] / res_w * 2 - 1, c[1] / res_w * 2 - res_h / res_w]).astype("float32")
    blk = np.concatenate([q, t, f, c, np.array(k, dtype=np.float64), np.array(p, dtype=np.float64)])
    return blk.astype(np.float32)


def generator_slot_scale(use_pre_angle: bool = True):
    """(half[37], mid[37]) of the generator's per-slot affine map (Fk_generator.py:143-168): slot i < 34 uses
    GAN_angle_range_table row 'joint{i+1}', slots 34-36 the global rotation table; with
    GAN_whether_use_preAngle off every slot is simply multiplied by 180."""
    if not use_pre_angle:
        return np.full(37, 180.0, np.float32), np.zeros(37, np.float32)
    rng = np.concatenate([GAN_ANGLE_RANGE, GAN_GLOBAL_ROT_RANGE]).astype(np.float64)
    return ((rng[:, 1] - rng[:, 0]) / 2).astype(np.float32), ((rng[:, 1] + rng[:, 0]) / 2).astype(np.float32)


def generator_src_col():
    """network column feeding each of the 37 slots (-1 for the slots forced to zero)."""
    src, col = [], 0
    for i in range(37):
        if i in GAN_ZERO_SLOTS:
            src.append(-1)
        else:
            src.append(col)
            col += 1
    return np.array(src, dtype=np.int64)
