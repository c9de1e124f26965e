import numpy as np, sys
sys.path.insert(0,'.')
import oracle as O
from birp_cpp_b200 import capi, synthetic as S
capi.set_device(0)
zyx, org, sp = S.make_cube(41, 37, 45, negative_slab=False)
W,H=27,16
poses=np.array([S.DEFAULT_POS,(30.,-25.,28.),(-14.,3.,5.),(1.5,-3.,2.)],np.float32)
fp=capi.frames_for_poses(poses,S.DEFAULT_AIM_X,27.,16.,W,H); fo=O.orc_frames(poses,S.DEFAULT_AIM_X,27.,16.,W,H)
c=capi.Cube(zyx,org,sp)
for tri in (False,True):
    for fl in (0,capi.NO_TEXTURE,capi.TEXTURE_ONLY,capi.GENERAL_LERP,capi.DIV_IEEE):
        got=c.render(fp,W,H,trilinear=tri,extra_flags=fl); want=O.orc_render_cube(zyx,org,sp,tri,fo,W,H)
        assert np.array_equal(got.view(np.uint32),want.view(np.uint32)),(tri,fl)
got=c.render(fp,W,H,nsamples=3000,trilinear=True)
P=capi.cmem_init(); capi.render_cmem(P,fp,W,H); capi.render_cmem(P,fp,W,H,fast=True)
capi.sample_cmem(P,np.random.rand(1000,3)*20-10); c.sample(np.random.rand(1000,3)*80-40,True)
capi.accumulate(np.random.rand(500,200).astype(np.float32),0.22); capi.ray_directions(fp[:1],W,H)
c.free(); print("sanitizer script ok")
