from .printer import print_model, source_hash
from .build import compile_model, compile_host_check, NVCC_FLAGS
