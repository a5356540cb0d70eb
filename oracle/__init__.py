"""CPU oracle for the DeeperCut forward path -- TEST INFRASTRUCTURE ONLY.

This package restates the reference's (eldar/deepcut-cnn, a BVLC-Caffe fork)
CPU algorithm for the hot path: Net::ForwardFromTo over the layer types the
deepercut deploy net instantiates.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product (``deepcut-cnn_b200``) never does; it fails loudly without its CUDA
library.

Pinning status: the layer functions are checked in ``tests/test_oracle_*.py``
against every known-answer / golden vector the reference's own tests hold for
this path: the naive ``caffe_conv`` reference incl.
dilation (test_convolution_layer.cpp:19-139), the Sobel identity (:498-590),
the deconvolution overlap test {3.1, 6.1, 12.1} (test_deconvolution_layer.cpp:
91-137), the GEMM integer answers (test_util_blas.cpp:20-88), the max-pool
literal matrices (test_pooling_layer.cpp:48-110), Scale/Bias broadcast, Eltwise
SUM, ReLU and Sigmoid properties.  Three behaviours are UNPINNED by the
reference's tests (BatchNorm use_global_stats, the custom CropLayer, the
whole-net output); for those the oracle is cross-checked against an
independent fp64 PyTorch model of the same prototxt (tests/test_oracle_net.py).
"""
