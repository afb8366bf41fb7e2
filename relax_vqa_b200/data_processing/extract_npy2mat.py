"""On-disk feature formats of the reference.

Per video the reference's drivers save one ``video_{i+1}_{feature_name}.npy`` holding the (T, D) per-frame
matrix (src/main_fragment_layerstack.py:348-354, src/main_layer_stack.py:198-203); ``extract_npy2mat.py``
then takes the temporal mean of each file and stacks the rows into an (N, D) ``.mat`` whose key is the dataset
name (src/data_processing/extract_npy2mat.py:62-85, 117-130).  Same names, same layout, so
``split_train_test.py`` / ``model_regression.py`` run unchanged on GPU-extracted features."""
import os

import numpy as np


def npy_name(index, network_name, compressed_type="original", resolution=None):
    """video_{i+1}_{network}_feature_map_original[_<resolution>].npy (0-based index in, 1-based in the name)."""
    feature_name = f"{network_name}_feature_map_original" + (f"_{resolution}" if resolution else "")
    return f"video_{index + 1}_{feature_name}.npy"


def features_dir(base_path, network_name, layer_name, data_name, compressed_type="original", resolution=None):
    """ref get_feature_path (:55-60)."""
    if data_name == "youtube_ugc":
        return f"{base_path}/{network_name}/{layer_name}/resolution_ugc/{compressed_type}_{resolution}/"
    return f"{base_path}/{network_name}/{layer_name}/{data_name}/{compressed_type}/"


def save_video_npy(out_dir, index, network_name, per_frame_matrix, resolution=None):
    os.makedirs(out_dir, exist_ok=True)
    path = os.path.join(out_dir, npy_name(index, network_name, resolution=resolution))
    np.save(path, np.asarray(per_frame_matrix, dtype=np.float32))
    return path


def collate(features_path, n_videos, network_name, resolution=None):
    """Temporal mean of every per-video file -> zero-initialised (N, D) float64 matrix (ref :117-126)."""
    matrix = None
    for i in range(n_videos):
        data = np.load(os.path.join(features_path, npy_name(i, network_name, resolution=resolution)))
        average_data = np.mean(data, axis=0)
        if matrix is None:
            matrix = np.zeros((n_videos,) + average_data.shape)
        matrix[i] = average_data
    return matrix


def save_features(mat_dir, data_name, matrix, network_name, compressed_type="original", resolution=None):
    """ref save_features (:62-85): <network>_<data>_<compressed>[_<res>]_features.mat with key = dataset name."""
    import scipy.io
    os.makedirs(mat_dir, exist_ok=True)
    if data_name == "youtube_ugc":
        name = os.path.join(mat_dir, f"{network_name}_{data_name}_{compressed_type}_{resolution}_features.mat")
        scipy.io.savemat(name, {f"{data_name}_{resolution}": matrix})
    else:
        name = os.path.join(mat_dir, f"{network_name}_{data_name}_{compressed_type}_features.mat")
        scipy.io.savemat(name, {data_name: matrix})
    return name
