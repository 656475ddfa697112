"""Drop-in for the reference's `cl` module (/root/reference/cl.py), served by liboclcg.so on B200.

The Python drivers do `import cl as pcl` and call (p_h-PY_C-CL.py:1933,1965,3609-3610;
p_h-PY_C-CL-multi-GPU.py:2136-2137,2161-2164,3662-3665; p_helmholtz.py:31,1839,1873):

    ctx, queue = pcl.initialize_cl_environment()                       cl.py:16-19
    ctx, queue = pcl.initialize_cl_environment_with_device(device)     cl.py:21-24
    devices    = pcl.get_gpu_devices()                                 cl.py:26-31
    kernels    = pcl.load_and_build_kernels(ctx, n_rhs)                cl.py:33-42
    kernels    = pcl.create_kernels(n_rhs)                             (older API, p_helmholtz.py:31)
    x = pcl.CG(ctx, queue, kernels, size, non_zeros, a_values, b_values,
               a_pointers, a_cols, x, n_rhs, n_iterations, device=None)   cl.py:44-200
    x = pcl.CG(size, non_zeros, a_values, b_values, a_pointers, a_cols,
               x, n_rhs, n_iterations)                                 (older API, p_helmholtz.py:1839)
    x = pcl.conjugate_gradient_multi_gpu(ctx, queue, kernels, ..., device)  cl.py:203-360

Same names, same argument meaning, same result convention (x is filled in place AND
returned, cl.py:188,200).  Contexts, queues and kernel dictionaries become light tokens:
there is nothing to JIT-compile, the sm_100a kernels are inside liboclcg.so, and the
"context" is just a CUDA device ordinal.  Put this directory on PYTHONPATH ahead of
the reference checkout and the drivers run unchanged.

Differences from the reference, all deliberate:
  - the value type follows the dtype of `a_values` (complex64 as the drivers pass,
    but also float32 / float64 / complex128) instead of the hard-coded IS_COMPLEX;
  - no recompilation when n_rhs == 1 (cl.py:45-46);
  - the matrix stays resident on the device between calls with the same content.
"""
import numpy as np

try:
    from . import _lib
except ImportError:  # `import cl` with this directory on sys.path, as the reference drivers do
    import _lib

IS_COMPLEX = True          # cl.py:5
WAVE_SIZE = 32             # cl.py:6
LOCAL_SIZE = 8 * WAVE_SIZE  # cl.py:7

KERNEL_NAMES = ("axpy", "aypx", "spmv", "sub", "vdot")   # cl.py:36-42


class Device:
    """Stands for a pyopencl Device: one CUDA device ordinal."""

    def __init__(self, ordinal):
        self.ordinal = int(ordinal)
        self.name = f"cuda:{self.ordinal}"

    def __repr__(self):
        return f"<cl.Device {self.name}>"

    def __hash__(self):
        return hash(self.ordinal)

    def __eq__(self, other):
        return isinstance(other, Device) and other.ordinal == self.ordinal


class Context:
    """Stands for a pyopencl Context."""

    def __init__(self, device):
        self.device = device
        self.devices = [device]


class CommandQueue:
    """Stands for a pyopencl CommandQueue; the engine owns one CUDA stream per resident matrix."""

    def __init__(self, ctx):
        self.context = ctx

    def flush(self):
        pass

    def finish(self):
        pass


def get_gpu_devices():
    """cl.py:26-31 -- every GPU of the box."""
    return [Device(i) for i in range(_lib.lib().cgb200_device_count())]


def initialize_cl_environment():
    """cl.py:16-19 -- a context on the default device."""
    devs = get_gpu_devices()
    if not devs:
        raise RuntimeError("no CUDA device: the B200 CG engine has no CPU fallback")
    ctx = Context(devs[0])
    return ctx, CommandQueue(ctx)


def initialize_cl_environment_with_device(device):
    """cl.py:21-24."""
    if not isinstance(device, Device):
        device = Device(device)
    ctx = Context(device)
    return ctx, CommandQueue(ctx)


def load_and_build_kernels(ctx, n_rhs):
    """cl.py:33-42 -- the kernels are precompiled; returns the same five keys."""
    _lib.lib()  # building / loading the library is the analogue of the OpenCL program build
    return {name: (name, int(n_rhs)) for name in KERNEL_NAMES}


def create_kernels(n_rhs):
    """Older module API used by p_helmholtz.py:31."""
    return load_and_build_kernels(None, n_rhs)


def _solve(device_ordinal, size, non_zeros, a_values, b_values, a_pointers, a_cols, x, n_rhs, n_iterations):
    a_values = np.ascontiguousarray(a_values)
    dt = a_values.dtype
    if dt not in _lib.DTYPE_CODE:
        raise TypeError(f"a_values dtype {dt} is not float32/float64/complex64/complex128")
    if not isinstance(x, np.ndarray) or x.dtype != dt or not x.flags["C_CONTIGUOUS"]:
        raise TypeError("x must be a C-contiguous numpy array with the dtype of a_values (written in place)")
    size, non_zeros, n_rhs, n_iterations = int(size), int(non_zeros), int(n_rhs), int(n_iterations)
    b_values = np.ascontiguousarray(b_values, dtype=dt)
    a_pointers = np.ascontiguousarray(a_pointers, dtype=np.intc)
    a_cols = np.ascontiguousarray(a_cols, dtype=np.intc)
    if a_values.size < non_zeros or a_cols.size < non_zeros or a_pointers.size < size + 1:
        raise ValueError("CSR arrays are shorter than size / non_zeros say")
    if b_values.size < size * n_rhs or x.size < size * n_rhs:
        raise ValueError("b_values / x must hold n_rhs blocks of `size` values")
    _lib.check(_lib.lib().cgb200_cg(device_ordinal, _lib.DTYPE_CODE[dt], size, non_zeros,
                                    _lib.ptr(a_values), _lib.ptr(b_values), _lib.ptr(a_pointers),
                                    _lib.ptr(a_cols), _lib.ptr(x), n_rhs, n_iterations))
    return x


def _ordinal(ctx, device):
    if device is not None:
        return device.ordinal if isinstance(device, Device) else int(device)
    if isinstance(ctx, Context):
        return ctx.device.ordinal
    return 0


def CG(*args, device=None):
    """cl.py:44-200.  Accepts the 12-argument form (ctx, queue, kernels, size, ...) and the
    older 9-argument form (size, ...) of p_helmholtz.py:1839; a 13th positional `device` too."""
    if len(args) in (12, 13):
        ctx = args[0]
        if len(args) == 13:
            device = args[12]
        return _solve(_ordinal(ctx, device), *args[3:12])
    if len(args) in (9, 10):
        if len(args) == 10:
            device = args[9]
        return _solve(_ordinal(None, device), *args[:9])
    raise TypeError("CG(ctx, queue, kernels, size, non_zeros, a_values, b_values, a_pointers, a_cols, x, "
                    "n_rhs, n_iterations[, device]) or CG(size, ..., n_iterations)")


def conjugate_gradient_multi_gpu(ctx, queue, kernels, size, non_zeros, a_values, b_values, a_pointers, a_cols,
                                 x, n_rhs, n_iterations, device):
    """cl.py:203-360 -- the same solve pinned to `device`; called from one thread per device
    by distribute_computations_with_threads (p_h-PY_C-CL-multi-GPU.py:2142-2181).  The C call
    releases the GIL and each device has its own resident matrix and stream."""
    return _solve(_ordinal(ctx, device), size, non_zeros, a_values, b_values, a_pointers, a_cols, x,
                  n_rhs, n_iterations)


def PCG(ctx, queue, kernels, size, non_zeros, a_values, b_values, a_pointers, a_cols, x, n_rhs, n_iterations,
        m_inv_diag=None, tol=0.0, device=None):
    """NOT in the reference's cl.py: the device twin of its numpy PCG (helmFE_var.py:546-586) in the calling
    convention of `CG` above, for the drivers that want a preconditioned subdomain solve (the report's own
    future work).  m_inv_diag: `size` values of M = inverse diagonal (None: 1 / diag(A), Jacobi);
    tol: absolute, on sqrt|r.r| as the reference stops (0: exactly n_iterations).  x is filled in place and returned."""
    try:
        from . import engine
    except ImportError:
        import engine
    a_values = np.ascontiguousarray(a_values)
    with engine.Matrix(a_values[:int(non_zeros)], np.asarray(a_pointers)[:int(size) + 1], np.asarray(a_cols)[:int(non_zeros)],
                       n=int(size), device=_ordinal(ctx, device)) as M:
        M.solve_pcg(b_values, x=x, k=int(n_rhs), M_inv_diag=m_inv_diag, max_iterations=int(n_iterations), tol=float(tol))
    return x
