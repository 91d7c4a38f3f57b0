"""pointnet_autoencoder_b200 -- B200-native (sm_100a) reconstruction-loss ops of
pointnet-autoencoder behind the reference's tf_ops Python API.

    from pointnet_autoencoder_b200.tf_ops.nn_distance import tf_nndistance
    from pointnet_autoencoder_b200.tf_ops.approxmatch import tf_approxmatch

(The directory is spelled with an underscore because a Python package name cannot
contain '-'.)  Importing the package does not load the CUDA library; the first op
call does, and raises if libpnae.so is missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
