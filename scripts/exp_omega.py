import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import test_bench_config_gpu as T
from helpers import rms_rel
from oracle import performer_oracle as PO
for snap in (0,1):
    m, sd, om, tok, seg = T._model(torch.bfloat16)
    if snap:
        c = 0.35355339059327373 * 1.4426950408889634
        om = ((om * c).to(torch.bfloat16).float() / c)
        m.fixed_omegas = om.cuda()
    ref_h=[]
    ref = T._oracle_rows(sd, om, tok, seg, [0, 73], ref_h)
    with torch.no_grad():
        hid,_ = m._forward_hidden(tok.cuda(), seg.cuda(), save=False)
    hid = hid.view(74, 2048, 512)
    got = torch.stack([hid[0], hid[73]]).float().cpu()
    print("snap", snap, "hidden rms rel", rms_rel(got, torch.cat(ref_h,0)))
