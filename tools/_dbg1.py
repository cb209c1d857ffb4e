import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np, anime4kcpp_b200 as A, oracle_lib as O
s = A.Session(0)
img = O.noise_u8(45, 61, 3, seed=32)
for name in ["acnet-legacy-hdn1"]:
    m = A.Model(name)
    got = s.process_host(m, img, 2.0); want = O.oracle_process(name, img, 2.0)
    d = np.abs(got.astype(int)-want.astype(int))
    print(name, "rgb", d.max(), (d==0).mean(), "per-channel exact", [(d[...,c]==0).mean() for c in range(3)])
    ys, xs = np.nonzero(d.max(axis=2)>0); print(len(ys), list(zip(ys[:10], xs[:10])))
    # isolate: y plane / uv
    y, uv = s.rgb2yuv(img)
    yo = np.empty((45,61),np.uint8); uvo = np.empty((45,61,2),np.uint8)
    O.oracle().orc_rgb2yuv(img.ctypes.data, 61,45,3,img.strides[0],O.U8, yo.ctypes.data, yo.strides[0], uvo.ctypes.data, uvo.strides[0])
    print("rgb2yuv exact", np.array_equal(y,yo), np.array_equal(uv,uvo))
    y2 = s.process_host(m, yo, 2.0); y2o = O.oracle_process(name, yo, 2.0); print("luma", O.compare_u8(y2,y2o))
    uv2 = s.resize_catmull_rom(uvo, 122, 90); uv2o = np.empty((90,122,2),np.uint8)
    O.oracle().orc_resize_catmull_rom(uvo.ctypes.data,61,45,2,uvo.strides[0],O.U8,uv2o.ctypes.data,122,90,uv2o.strides[0]); print("resize", O.compare_u8(uv2,uv2o))
    rgb2 = s.yuv2rgb(y2o, uv2o); back = np.empty((90,122,3),np.uint8)
    O.oracle().orc_yuv2rgb(y2o.ctypes.data,y2o.strides[0],uv2o.ctypes.data,uv2o.strides[0],122,90,3,O.U8,back.ctypes.data,back.strides[0]); print("merge", O.compare_u8(rgb2,back))
    print("final vs oracle merge of oracle parts", O.compare_u8(got, back))
g = O.noise_u8(40,48,1,seed=1234).astype(np.uint16)*257
got = s.process_host(A.Model("acnet-legacy-hdn0"), g, 2.0); want = O.oracle_process("acnet-legacy-hdn0", g, 2.0)
d = np.abs(got.astype(int)-want.astype(int)); print("u16", d.max(), (d==0).mean(), (d<=1).mean())
