"""saspa_aug_b200 -- B200-native (sm_100a) implementation of SaSPA's augmentation-generation hot path.

Importable name of the package the task calls ``saspa-aug_b200`` (a symlink of that name sits beside it).
See DESIGN.md for the path, the boundary and the kernels; include/saspa_b200.h for the C ABI."""
__version__ = "0.1.0"
