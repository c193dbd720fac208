from .printer import print_model, source_hash
from .build import compile_model, NVCC_FLAGS
