"""The reference's pose text files for users of the batched Python binding: same columns and number
formatting as the C++ writers in host/PoseFiles.cpp (default std::ostream formatting of a float == printf("%g")).

    poses_orig.txt          src/main.cpp:373               frameId kfId poseWrtWorld[6] rescaleFactor depthMapOccupancy
    matchframes*.txt        src/main.cpp:382, src/GlobalOptimize.cpp:580   frameId kfId poseWrtOrigin[6] rescale seeds matchValue rms angle
"""
import numpy as np


def _g(x):
    return "%g" % float(np.float32(x))


def orig_pose_line(frame_id, kf_id, pose_wrt_world, rescale_factor, occupancy, batch_start_id=0):
    cols = [str(int(frame_id) + batch_start_id - 1), str(int(kf_id) + batch_start_id - 1)] + [_g(v) for v in pose_wrt_world]
    return " ".join(cols + [_g(rescale_factor), _g(occupancy)]) + "\n"


def match_pose_line(frame_id, kf_id, pose_wrt_origin, rescale_factor, seeds, match_value=None, rms_error=None, view_angle=None,
                    batch_start_id=0):
    """match_value None => the sequential-pair form of src/main.cpp:382 (literal "0 0 0" tail)."""
    cols = [str(int(frame_id) + batch_start_id - 1), str(int(kf_id) + batch_start_id - 1)] + [_g(v) for v in pose_wrt_origin]
    tail = ["0", "0", "0"] if match_value is None else [_g(match_value), _g(rms_error), _g(view_angle)]
    return " ".join(cols + [_g(rescale_factor), _g(seeds)] + tail) + "\n"


def read_pose_file(path):
    """Rows of poses_orig.txt / matchframes*.txt as a float64 array (ids in columns 0, 1)."""
    rows = [l.split() for l in open(path) if l.strip()]
    return np.array(rows, np.float64)


def chain_world_poses(concat_relative, rel_poses, kf_world):
    """poseWrtWorld = log(exp(rel) exp(kf_world)) (frame::calculatePoseWrtWorld, src/Frame.cpp:352-372) for a list of relative
    poses on one keyframe; `concat_relative` is capi.concat_relative (device algebra) or the oracle's."""
    return [concat_relative(np.asarray(r, np.float32), np.asarray(kf_world, np.float32)) for r in rel_poses]
