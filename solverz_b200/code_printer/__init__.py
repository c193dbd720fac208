from .make_module import module_printer, made_numerical, render_cuda_module
