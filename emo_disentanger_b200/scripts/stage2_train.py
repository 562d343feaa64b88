"""Stage-2 training -- the reference's `python3 stage2_accompaniment/train.py -m {performer,gpt2} -c CONFIG
-r {remi,functional}` surface (argparse flags, YAML schema, log.txt / valloss.txt formats, checkpoint file
names and state-dict keys: train.py:21-29,198-356) over the B200 hot path.

    python -m emo_disentanger_b200.scripts.stage2_train -m performer -c CONFIG.yaml -r functional
    torchrun --nproc-per-node 8 -m emo_disentanger_b200.scripts.stage2_train ...      (data parallel)

Differences by design: the epoch loop calls the fused `train_step` (no autograd graph, no [B,T,V] logits
copied to the host: loss / accuracy come back as 3 floats) and `FusedAdam` (clip 0.5 + Adam in one launch);
under torchrun the batches are rank-strided and gradients all-reduced once per optimiser step; rank 0
alone logs and checkpoints.  Datasets stay the reference's (imported from the reference tree) unless the
config has a `synthetic` data section or sets `data_loader.gpu_token_store: true` (pieces resident in HBM, batches
assembled on the device: data/token_store.py)."""
import argparse
import os
import shutil
import time
import numpy as np
import torch
import yaml

from .. import dp
from ..optim import FusedAdam, WarmupCosine
from ..stage2 import MusicPerformer, MusicGPT2
from . import common


def build_model(model_type, vocab_size, mc, gpuid, compute_dtype=torch.bfloat16):
    kw = dict(use_segment_emb=mc['use_segemb'], n_segment_types=mc.get('n_segment_types', 2),
              use_chord_mhot_emb=False, compute_dtype=compute_dtype)
    if model_type == "performer":
        m = MusicPerformer(vocab_size, mc['n_layer'], mc['n_head'], mc['d_model'], mc['d_ff'], mc['d_embed'],
                           favor_feature_dims=mc['feature_map']['n_dims'], **kw)
    elif model_type == "gpt2":
        m = MusicGPT2(vocab_size, mc['n_layer'], mc['n_head'], mc['d_model'], mc['d_ff'], mc['d_embed'], **kw)
    else:
        raise NotImplementedError("Unsuppported model: %s" % model_type)
    return m.cuda(gpuid)


def load_params(model, path):
    pre = torch.load(path, map_location='cpu')
    pre = {k: v for k, v in pre.items() if 'feature_map.omega' not in k}          # train.py:306-308
    sd = model.state_dict()
    sd.update(pre)
    model.load_state_dict(sd)


class Log:
    def __init__(self, ckpt_dir, enabled):
        self.path, self.enabled = os.path.join(ckpt_dir, 'log.txt'), enabled

    def write(self, ep, steps, loss, secs):
        if not self.enabled:
            return
        if not os.path.exists(self.path):
            with open(self.path, 'w') as f:
                f.write('{:4} {:8} {:12} {:12}\n'.format('ep', 'steps', 'recons_loss', 'ep_time'))
        with open(self.path, 'a') as f:
            f.write('{:<4} {:<8} {:<12} {:<12}\n'.format(ep, steps, round(loss, 5), round(secs, 2)))


def main(argv=None):
    ap = argparse.ArgumentParser(description='')
    req = ap.add_argument_group('required arguments')
    req.add_argument('-m', '--model_type', choices=['performer', 'gpt2'], required=True, help='model backbone')
    req.add_argument('-c', '--configuration', required=True, help='configurations of training')
    req.add_argument('-r', '--representation', choices=['remi', 'functional'], required=True,
                     help='representation for symbolic music')
    ap.add_argument('--fp32', action='store_true', help='fp32 parity mode (SIMT kernels)')
    ap.add_argument('--max_steps', type=int, default=None, help='stop after this many micro-batches (smoke runs)')
    args = ap.parse_args(argv)
    conf = yaml.load(open(args.configuration), Loader=yaml.FullLoader)
    tc, dc, mc = conf['training'], conf['data_loader'], conf['model']
    rank, local, world = dp.init_from_env()
    gpuid = local if world > 1 else tc['gpuid']
    torch.cuda.set_device(gpuid)
    rep = args.representation
    warmup, max_lr, min_lr = tc['warmup_steps'], tc['lr'], tc['lr_scheduler']['eta_min']
    ckpt_dir = tc['ckpt_dir'].format(rep)
    accum = tc.get('accum_steps', 1)
    log_interval, ckpt_interval = tc['log_interval'], tc['ckpt_interval']

    if dc.get('synthetic'):
        sy = dc['synthetic']
        T = sy.get('seq_len', 2048)
        dset = common.SyntheticStage2(sy['vocab_size'], dc['batch_size'], T, sy['n_train_batches'], 0)
        vset = common.SyntheticStage2(sy['vocab_size'], dc['batch_size'], T, sy['n_val_batches'], 10 ** 6)
        dloader, vloader = dset, vset
    elif dc.get('gpu_token_store'):
        # pieces tokenised once into HBM, batches assembled by one kernel launch (data/token_store.py): same batch
        # dicts as the reference Dataset + DataLoader, same start-bar choice (`random.choice` over the admissible bars)
        import pickle
        from ..data import Stage2TokenStore
        ddir, vfile = dc['data_path'].format(rep), dc['vocab_path'].format(rep)
        mk = lambda split: Stage2TokenStore.from_files([os.path.join(ddir, p_) for p_ in pickle.load(open(split, 'rb'))], vfile,
                                                       model_dec_seqlen=mc['max_len'], predict_key=False, device='cuda:%d' % gpuid)
        dset, vset = mk(dc['train_split']), mk(dc['val_split'])

        dloader = common.StoreEpochs(dset, dc['batch_size'], True, rank, world, seed=dc.get('seed', 0))
        vloader = common.StoreEpochs(vset, dc['batch_size'], True)                 # rank 0 validates alone
    else:
        from torch.utils.data import DataLoader
        dl = common.reference_module('stage2_accompaniment', 'dataloader')
        ut = common.reference_module('stage2_accompaniment', 'utils')
        mk = lambda split: dl.REMISkylineToMidiTransformerDataset(
            data_dir=dc['data_path'].format(rep), vocab_file=dc['vocab_path'].format(rep),
            model_dec_seqlen=mc['max_len'], pieces=ut.pickle_load(split), pad_to_same=True, predict_key=False)
        dset, vset = mk(dc['train_split']), mk(dc['val_split'])
        dloader = DataLoader(dset, batch_size=dc['batch_size'], shuffle=True, num_workers=8,
                             generator=common.shared_generator(dc.get('seed', 0)))   # same permutation on every rank
        vloader = DataLoader(vset, batch_size=dc['batch_size'], shuffle=True, num_workers=8,
                             generator=common.shared_generator(1))

    torch.manual_seed(0)
    model = build_model(args.model_type, dset.vocab_size, mc, gpuid, torch.float32 if args.fp32 else torch.bfloat16)
    if tc.get('trained_params'):
        load_params(model, tc['trained_params'])
    model.train()
    sync = dp.GradSync(model)
    sync.broadcast_params()
    print('# params:', sum(p.numel() for p in model.parameters() if p.requires_grad))
    print('segemb:', model.segemb)
    opt = FusedAdam(model, lr=max_lr, max_grad_norm=0.5, grad_sync=sync)
    sched = WarmupCosine(opt, max_lr, min_lr, warmup, tc['lr_scheduler']['T_max'], accum)
    if tc.get('trained_optim'):
        opt.load_state_dict(torch.load(tc['trained_optim'], map_location='cpu'))
    params_dir, optim_dir = os.path.join(ckpt_dir, 'params/'), os.path.join(ckpt_dir, 'optim/')
    if rank == 0:
        for d_ in (ckpt_dir, params_dir, optim_dir):
            os.makedirs(d_, exist_ok=True)
        shutil.copy(args.configuration, os.path.join(ckpt_dir, 'config.yaml'))
    log = Log(ckpt_dir, rank == 0)

    steps = 0
    for ep in range(1, tc['num_epochs'] + 1):
        model.train()
        st = time.time()
        stats = torch.zeros(3, device='cuda')            # epoch sums of [count, loss_sum, n_correct]
        n_tok = 0
        for batch in common.rank_strided(dloader, rank, world):
            x, tgt, seg = (batch[k].cuda(gpuid, non_blocking=True) for k in ('dec_input', 'dec_target', 'track_mask'))
            steps += 1
            sync.begin_step(last_micro_batch=(steps % accum == 0))    # per-layer gradient buckets leave from the last backward only
            acc = model.train_step(x, seg, tgt, gscale=1.0 / accum, count_allreduce=sync.count_allreduce)
            if steps % accum == 0:
                opt.step()                               # all-reduce -> clip(0.5) -> Adam -> zero grads
            sched.update(steps)
            stats += sync.allreduce_stats(acc) if world > 1 else acc
            n_tok += x.numel() * world
            if steps % log_interval == 0:
                s = stats.tolist()
                log.write(ep, steps, s[1] / max(s[0], 1.0), time.time() - st)
            if args.max_steps and steps >= args.max_steps:
                break
        s = stats.tolist()
        loss = s[1] / max(s[0], 1.0)
        secs = time.time() - st
        if rank == 0:
            print('[epoch {:03d}] training completed\n  -- loss = {:.4f}\n  -- time elapsed = {:.2f} secs.  '
                  '({:.0f} tokens/s)'.format(ep, loss, secs, n_tok / secs))
        log.write(ep, steps, loss, secs)
        if rank == 0 and ep % ckpt_interval == 0:
            torch.save(model.state_dict(), os.path.join(params_dir, 'ep{:03d}_loss{:.3f}_params.pt'.format(ep, loss)))
            torch.save(opt.state_dict(), os.path.join(optim_dir, 'ep{:03d}_loss{:.3f}_optim.pt'.format(ep, loss)))
        if rank == 0:
            vl, ta, ca, ma, oa = validate(model, vloader, vset.pad_token, gpuid)
            with open(os.path.join(ckpt_dir, 'valloss.txt'), 'a') as f:
                f.write("ep{:03d} | loss: {:.3f} | valloss: {:.3f} (±{:.3f}) | total_acc: {:.3f} | "
                        "chord_acc: {:.3f} | melody_acc: {:.3f} | others_acc: {:.3f}\n".format(
                            ep, loss, np.mean(vl), np.std(vl), np.mean(ta), np.mean(ca), np.mean(ma), np.mean(oa)))
        if args.max_steps and steps >= args.max_steps:
            break
    return 0


def validate(model, loader, pad_token, gpuid):
    """per-batch loss + total / chord / melody / others accuracy (train.py:121-193), argmax on the device"""
    model.eval()
    vl, ta, ca, ma, oa = [], [], [], [], []
    with torch.no_grad():
        for batch in loader:
            x, tgt, seg = (batch[k].cuda(gpuid) for k in ('dec_input', 'dec_target', 'track_mask'))
            ch, me = batch['chord_idx'].cuda(gpuid) == 1, batch['melody_idx'].cuda(gpuid) == 1
            logits = model(x, seg_inp=seg)
            vl.append(float(model.compute_loss(logits, tgt)['recons_loss']))
            hit = logits.argmax(-1) == tgt
            valid = tgt != pad_token
            nv, nc, nm = int(valid.sum()), int(ch.sum()), int(me.sum())
            tot = float(hit[valid].float().mean()) if nv else float('nan')
            c = float(hit[ch].float().mean()) if nc else float('nan')
            m_ = float(hit[me].float().mean()) if nm else float('nan')
            rest = nv - nc - nm
            o = (tot * nv - (c * nc if nc else 0) - (m_ * nm if nm else 0)) / rest if rest else float('nan')
            ta.append(tot); ca.append(c); ma.append(m_); oa.append(o)
    return vl, ta, ca, ma, oa


if __name__ == "__main__":
    raise SystemExit(main())
