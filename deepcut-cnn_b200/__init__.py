"""deepcut-cnn_b200: B200-native DeeperCut forward path.

Layout
  csrc/            hand-written sm_100a CUDA + the extern "C" ABI (libdeepcut_b200.so)
  caffe_host/      C++ host keeping Caffe's Net/Layer/Blob API over that ABI (libcaffe_b200.so)
  libdc.py         ctypes binding of include/deepcut_b200.h (tests / bench / harness)
  gen_prototxt.py  emits the deploy prototxt; synth.py seeded inputs + trained-like weights
  build.py         nvcc / g++ recipes (in-tree .so files; they travel to the GPU box)

The directory name is not a Python identifier; import it with
``importlib.import_module("deepcut-cnn_b200")``.
"""
