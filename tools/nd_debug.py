import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mfs_b200.multi_dims.multi_indices import generate_graded_lexico_multi_indices as gen, gram_and_hankel_indices_graded_lexico as gh
from mfs_b200.multi_dims.filtering import moment_filter_nd_rms, moment_filter_nd_cms
from mfs_b200.multi_dims.moments import sde_cond_moments_euler_maruyama
from mfs_b200.multi_dims.ss_models import prey_predator
from oracle import mfs_oracle_nd as ND
N,B,T=5,6,12
mis=gen(2,2*N-1); inds=gh(N,2)
dt,_,ts,gs,drift,dispersion,emission,pmf,simulate=prey_predator(mis)
rng=np.random.Generator(np.random.PCG64(675))
_,xs,ys=simulate(rng,integration_steps=10,T=T,n=B)
fam=sde_cond_moments_euler_maruyama(drift,dispersion,dt,mis)
cmss,means,nell,status=moment_filter_nd_cms((fam[1],'index'),fam[3],pmf,torch.from_numpy(ys).cuda(),(mis,inds),gs.cms,gs.mean,return_status=True)
print('gpu status',status.cpu().numpy())
f_r,f_c,f_m=ND.lv_cond_moments('euler',mis,use_kan=False)
for k in range(B):
    rc,rm,rn=ND.moment_filter_nd_cms(f_c,f_m,ND.lv_measurement_pmf,ys[k],(mis,inds),gs.cms,gs.mean)
    bad=np.argmax(~np.isfinite(rm[:,0])) if not np.isfinite(rn) else -1
    fin=np.isfinite(rm[:,0]) & np.isfinite(means[k,:,0].cpu().numpy())
    print(k,'oracle first bad',bad,'mean err',np.max(np.abs(means[k].cpu().numpy()[fin]-rm[fin])) if fin.any() else None,'nell',nell[k].item(),rn)
# timing at scale
B2=20000; T2=50
ys2=torch.from_numpy(np.tile(ys[:1,:],(B2,5))[:, :T2].copy()).cuda()
for it in range(2):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); out=moment_filter_nd_cms((fam[1],'index'),fam[3],pmf,ys2,(mis,inds),gs.cms,gs.mean,history='last',return_status=True); e1.record(); torch.cuda.synchronize()
print('N=5 nd filter', B2*T2/e0.elapsed_time(e1)*1e3,'steps/s', 'diverged',(out[-1]>=0).double().mean().item())
