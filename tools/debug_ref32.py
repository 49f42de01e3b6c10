import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from oracle import boxes as ob, pk_weights
from saclaymocks_b200.boxes import BoxSynth, WEIGHT_OF
from saclaymocks_b200 import spectra as sp
NX,NY,NZ,dcell=32,32,1536,2.19
W=pk_weights.weights(NX,NY,NZ,dcell)
noise=ob.draw_noise(NX,NY,NZ,42)
raw,p0,boxes,sig=ob.make_boxes(NX,NY,NZ,dcell,42,W,noise=noise,products=("box","eta_xx"))
cuda=torch.device("cuda:0")
bs=BoxSynth(NX,NY,NZ,dcell,device=cuda)
boxk=bs.draw_grf_boxk(noise=torch.as_tensor(noise,device=cuda))
gk=bs.boxk_to_numpy(boxk)
print("boxk rel", np.sqrt((np.abs(gk-raw)**2).sum()/(np.abs(raw)**2).sum()), "max abs", np.abs(gk-raw).max(), np.abs(raw).max())
i=np.unravel_index(np.argmax(np.abs(gk-raw)),raw.shape); print("argmax",i,gk[i],raw[i])
Wd={k:bs.upload_weights(v) for k,v in W.items()}
box,_=bs.synth(boxk,"box",wtable=Wd["P0"])
b=box.cpu().numpy(); r=boxes["box"]
print("box rel",np.sqrt(((b-r)**2).sum()/(r**2).sum()),"maxabs",np.abs(b-r).max())
d=np.fft.rfftn((b-r).astype(np.float64)); rk=np.fft.rfftn(r.astype(np.float64))
for ax in range(3):
    other=tuple(a for a in range(3) if a!=ax)
    e=(np.abs(d)**2).sum(axis=other); s=(np.abs(rk)**2).sum(axis=other)
    rel=np.sqrt(e/s); j=np.argsort(rel)[-5:]
    print("axis",ax,"worst modes",j,rel[j])
print("mean diff", (b-r).mean(), r.mean(), b.mean())
# ---- gather check on identical (oracle) boxes
from oracle import spectra as osp
from helpers import qso_files_from_golden
g=dict(np.load("tests/golden/ref_ref32.npz"))
_,_,allb,_=ob.make_boxes(NX,NY,NZ,dcell,42,W,noise=noise)
geom=sp.SkewerGeometry(NX,NY,NZ,dcell)
q=np.concatenate(qso_files_from_golden(g))
xyzr,nfor=sp.qso_lines_of_sight(geom,q["RA"],q["DEC"],q["Z_QSO_RSD"],190.0,0.0)
eng=sp.SkewerEngine(geom,device=cuda)
f={k:torch.as_tensor(allb[k],device=cuda) for k in sp.FIELDS}
out=eng.read_spec(f,xyzr,np.maximum(nfor,0))
d=out[0].cpu().numpy()
og=osp.Geometry(NX,NY,NZ,dcell)
qf=qso_files_from_golden(g)
lam32=np.float32(geom.lambda_vec); ids=list(q["THING_ID"])
for s in range(4):
    for p in osp.make_spectra_slice(og,allb,qf,s,4,190.0,0.0):
        idx=np.searchsorted(lam32,p["lam"]); row=ids.index(p["id"])
        m=p["delta_l"]>-1e5
        if m.any():
            err=np.abs(d[row,idx]-p["delta_l"])[m]
            print(s,p["id"],"n",m.sum(),"maxerr",err.max(),"at",err.argmax(),"mean",err.mean(), "X/R", xyzr[row,0]/xyzr[row,3], xyzr[row,1]/xyzr[row,3])
