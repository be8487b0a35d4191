"""jlama_b200: Blackwell (sm_100a) back-end for tjake/Jlama's quantized forward pass.

The product is the shared library libjlama_b200.so (C ABI: include/jlama_b200.h) -- hand-written
CUDA kernels plus a C++ host driver.  The Python modules here are the host-side mirror of the
reference's plug-in interface used by tests and benchmarks:

  native   ctypes binding (stand-in for the Panama-FFI binding)
  tensor   AbstractTensor types and block quantisers (core/tensor/*)
  ops      CudaTensorOperations  (core/tensor/operations/TensorOperations.java)
  model    LlamaModel / DistributedContext (core/model/*)
  synth    public model dims + seeded synthetic checkpoints

Nothing in this package imports oracle/ and nothing falls back to the CPU.
"""
__version__ = "0.1.0"
