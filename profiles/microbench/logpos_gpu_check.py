import sys, numpy as np, torch
sys.path.insert(0, __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.abspath(__file__)), "..", ".."))
from tests import golden_util as gu
from collaborative_nonstationary_multivariate_gaussian_process_b200 import logpos
g = gu.load("sim_logpos"); dev = "cuda:0"
d = lambda k: torch.from_numpy(np.asarray(g[k], dtype=np.float64)).to(dev); sc = lambda v: torch.tensor(float(v), dtype=torch.float64, device=dev)
hyp=[sc(v) for v in g["hyp"]]; a,b,c=(float(v) for v in g["abc"]); ts2=sc(g["ts2"])
def rel(out, ref): return max(abs(float(o)-r)/abs(r) for o,r in zip(out, np.atleast_1d(ref)))
print("logpos", rel(logpos.logpos(d("tilde_l"), d("tilde_sigma"), d("uL_vec"), ts2, d("Y"), d("x"), *hyp, a, b, c, verbose=True), g["logpos_verbose"]))
print("deviance", rel([logpos.deviance(d("tilde_l"), d("tilde_sigma"), d("L_vec"), ts2, d("Y"), d("x"))], g["deviance"]))
print("logpos_S", rel(logpos.logpos_S(sc(g["tlS"]), sc(g["tsS"]), d("uL_vec"), ts2, d("Y"), d("x"), sc(-1.0), sc(0.7), a, b, c, verbose=True), g["logpos_S_verbose"]))
ih=torch.from_numpy(g["ih"]).to(dev)
print("hadamard", rel(logpos.logpos_hadamard(d("tlh"), d("tsh"), d("L_vec"), ts2, d("xh"), ih, d("yh"), *hyp, a, b, c, verbose=True), g["logpos_hadamard_verbose"]))
print("hadamard_S", rel(logpos.logpos_hadamard_S(sc(g["tlS"]), sc(g["tsS"]), d("L_vec"), ts2, d("xh"), ih, d("yh"), sc(-1.0), sc(0.7), a, b, c, verbose=True), g["logpos_hadamard_S_verbose"]))
parsH = torch.cat([d("tlh"), d("tsh"), d("L_vec"), ts2.view(1)])
print("nlogpos_obj_hadamard", rel([logpos.nlogpos_obj_hadamard(parsH, d("xh"), ih, d("yh"), *[float(h) for h in hyp], a, b, c)], g["nlogpos_obj_hadamard"]))
hyp_i = [sc(v) for v in g["hyp_i"]]
print("logpos_SVC", rel(logpos.logpos_SVC(d("tli"), d("uLi"), ts2, d("Yi"), d("xi"), *hyp_i, a, b, verbose=True), g["logpos_SVC_verbose"]))
print("logpos_hadamard_SVC", rel(logpos.logpos_hadamard_SVC(d("tlh"), d("Lv_h"), ts2, d("xh"), ih, d("yh"), *hyp_i, a, b, verbose=True), g["logpos_hadamard_SVC_verbose"]))
