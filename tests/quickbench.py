import sys, time, importlib, os
sys.path.insert(0,'.'); sys.path.insert(0,'tests'); sys.path.insert(0,'deepcut-cnn_b200/python')
import numpy as np
import caffe, dcutil, netutil
libdc = dcutil.libdc
caffe.set_mode_gpu(); caffe.set_device(0)
n,h,w = [int(v) for v in sys.argv[1:4]]
path = '/tmp/b.prototxt'; dcutil.gen_prototxt.write(path, height=h, width=w)
net = caffe.Net(path, caffe.TEST)
rng = np.random.default_rng(0)
for name, blobs in net.params.items():
    for b in blobs:
        if 'bn' in name and tuple(b.shape)==(1,): b.data[...] = 1
        elif name.startswith('bn') : b.data[...] = rng.uniform(0.5,1.5,b.shape)
        else: b.data[...] = (rng.standard_normal(b.shape)*0.05).astype(np.float32)
net.blobs['data'].reshape(n,3,h,w)
net.blobs['data'].data[...] = dcutil.synth.images(n,h,w)
t=time.time(); net.forward(); caffe.sync(); print("first forward (plan+pack) s", time.time()-t, "fused", net.fused_last_forward, net.fusion_diagnostic, "launches", net.last_forward_launches)
for i in range(3): net.forward()
caffe.sync()
t=time.time()
K=5
for i in range(K): net.forward()
caffe.sync()
dt=(time.time()-t)/K
print("forward ms %.2f  images/s %.1f" % (dt*1e3, n/dt))
