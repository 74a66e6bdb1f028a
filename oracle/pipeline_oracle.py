"""CPU restatement of the reference's input pipeline (TEST INFRASTRUCTURE ONLY, like the rest of oracle/).

Follows, with the random draws turned into arguments so that the result is a function of its inputs:
  data_utils.py:162-168   center_data        data_utils.py:133-143   normalize_data
  data_utils.py:171-186   get_current_data_h5 (point subset idx_pts[:num_points], shared by all clouds)
  pointnet2/utils/provider.py:34-52 rotate_point_cloud   :189-200 jitter_point_cloud   :202-227 shift / scale   :229-236 dropout
numpy dtypes as in the reference: float32 clouds, float64 rotation matrix and jitter (results stored / fed as float32).
Parity unpinned: the reference holds no vectors for these functions (they consume np.random); the functions below are the
reference's own numpy expressions, so pinning reduces to numpy itself."""
import numpy as np


def center_data(pcs):
    pcs = np.array(pcs, dtype=np.float32, copy=True)
    for pc in pcs:
        centroid = np.mean(pc, axis=0)
        pc[:, 0] -= centroid[0]
        pc[:, 1] -= centroid[1]
        pc[:, 2] -= centroid[2]
    return pcs


def normalize_data(pcs):
    pcs = np.array(pcs, dtype=np.float32, copy=True)
    for pc in pcs:
        d = max(np.sum(np.abs(pc) ** 2, axis=-1) ** (1. / 2))
        pc /= d
    return pcs


def rotate_point_cloud(batch_data, angles):
    rotated = np.zeros(batch_data.shape, dtype=np.float32)
    for k in range(batch_data.shape[0]):
        cosval, sinval = np.cos(angles[k]), np.sin(angles[k])
        rotation_matrix = np.array([[cosval, 0, sinval], [0, 1, 0], [-sinval, 0, cosval]])
        rotated[k, ...] = np.dot(batch_data[k, ...].reshape((-1, 3)), rotation_matrix)
    return rotated


def jitter_point_cloud(batch_data, noise, sigma=0.01, clip=0.05):
    assert clip > 0
    jittered = np.clip(sigma * np.asarray(noise, np.float64), -1 * clip, clip)
    jittered += batch_data
    return jittered.astype(np.float32)          # what feed_dict does with the float64 array


def augment(src, n, perm=None, angles=None, scale=None, shift=None, noise=None, sigma=0.01, clip=0.05, drop=None,
            center=False, normalize=False):
    x = np.asarray(src, np.float32)
    if center:
        x = center_data(x)
    if normalize:
        x = normalize_data(x)
    idx = np.arange(n) if perm is None else np.asarray(perm)[:n]
    x = x[:, idx, :].copy()
    if drop is not None:
        for b in range(x.shape[0]):
            di = np.where(np.asarray(drop[b]).astype(bool))[0]
            if len(di) > 0:
                x[b, di, :] = x[b, 0, :]
    if angles is not None:
        x = rotate_point_cloud(x, np.asarray(angles, np.float64))
    if scale is not None:
        # provider.py:222-226 `batch_data[b] *= scales[b]` with a float64 scalar: numpy >= 2 (the one that runs the reference here)
        # multiplies in float64 and rounds to float32; numpy 1.x demoted the scalar and multiplied in float32 (<= 1 ulp apart)
        for b in range(x.shape[0]):
            x[b, :, :] *= np.float64(scale[b])
    if shift is not None:
        # provider.py:210-213 `batch_data[b] += shifts[b, :]`: float32 array += float64 array -> float64 add, rounded
        for b in range(x.shape[0]):
            x[b, :, :] += np.asarray(shift[b], np.float64)
    if noise is not None:
        x = jitter_point_cloud(x, noise, sigma, clip)
    return x
