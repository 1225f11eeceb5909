import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import torch_ref
import test_trainer_gpu as T
from action_conditioned_gans_b200.trainer import Trainer
worst = {}
for dna, loss, opt in [(True, "bce", "adam"), (True, "wass", "rmsprop"), (False, "bce", "adam")]:
    for rep in range(6):
        params = T._params(dna, 6)
        ora = torch_ref.Trainer(params, True, loss, opt, dna, ksize=6)
        trn = Trainer(None, True, loss, opt, dna, batch_size=8, ksize=6, params=params, precision="bf16")
        for it in range(3):
            img, nxt, act, state = T._feeds(8, 10 + it)
            if it == 0:
                gl = trn.pretrain_g(img, nxt, act, state); T._gates(trn); gl_ref = ora.pretrain_g(img, nxt, act, state)
                e = abs(gl - gl_ref) / abs(gl_ref); k = (dna, loss, "pretrain")
                worst[k] = max(worst.get(k, 0), e); continue
            s = trn.train_d(img, nxt, act, summarize=True); T._gates(trn); s_ref = ora.train_d(img, nxt, act, summarize=True)
            for key in ("discriminator_direct_loss", "discriminator_gen_loss", "discriminator_loss", "g_loss", "g_l2_loss"):
                e = abs(s[key] - s_ref[key]) / max(1.0, abs(s_ref[key])); k = (dna, loss, it, "d", key); worst[k] = max(worst.get(k, 0), e)
            fr = trn.train_g(img, nxt, act, state); T._gates(trn); fr_ref = ora.train_g(img, nxt, act, state)
            sg, sg_ref = trn.summaries(), ora.summaries()
            for key in ("g_loss", "g_l2_loss", "g_adv_loss", "g_psnr"):
                e = abs(sg[key] - sg_ref[key]) / max(1.0, abs(sg_ref[key])); k = (dna, loss, it, "g", key); worst[k] = max(worst.get(k, 0), e)
for k, v in sorted(worst.items(), key=lambda kv: -kv[1])[:12]:
    print("%.4f  %s" % (v, k))
