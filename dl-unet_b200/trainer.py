"""Drop-in for the reference's `trainer.py`: `training(unet, train_loader, val_loader, epochs,
batch_size, device, fold_dir, DATASET)` with the same loaders, prints, progress files and
checkpoints, whose hot loop (reference trainer.py:52-80 and 95-121) runs on the B200 kernels:

    forward            -> network.Unet.forward (sm_100a kernels)
    crop + one-hot + nn.BCEWithLogitsLoss(weight)  -> one fused loss kernel (fwd + dlogits)
    loss.backward()    -> explicit dgrad / wgrad schedule
    optim.SGD.step()   -> one fused kernel over the flat parameter arena

Deviation kept on purpose: the weight map is applied per pixel to both channels,
which is what the reference computes at batch size 1; its [B,H,W]-vs-[B,2,H,W] broadcast at B = 2
mixes samples and at B >= 3 raises.

Reference: trainer.py:15-235.
"""
import os
from time import time

import numpy as np
import torch
import torch.optim as optim

from b200.checkpoint import AsyncCheckpointer
from b200.train import FusedSGD, TrainStep, weighted_bce_with_logits
from functions import class_balance, evaluation_metrics, weighted_map

# The reference selects its stop criterion and weight-map generator with `DATASET is '<literal>'`
# (trainer.py:18-26, 68, 110): an identity test that only holds for the very same string object, so
# a name coming from argparse never matches.  The identity semantics are kept, against these objects.
HELA, ISBI, U373 = "DIC-C2DH-HeLa", "ISBI2012", "PhC-C2DH-U373"
_GOALS = ((HELA, 0.7756), (ISBI, 0.0611), (U373, 0.9203))  # trainer.py:20,23,26


def _goal_for(dataset):
    for name, goal in _GOALS:
        if dataset is name:
            return goal
    return None


def _weights_for(dataset, labels):
    """trainer.py:68-73 -- weight map for the batch ([B,L,L] fp32).  `labels` already lives on the training
    device: on CUDA the maps are built there (b200/weightmap.py) instead of by OpenCV on the host."""
    gt = labels.squeeze(1)
    return weighted_map(gt) if dataset is HELA else class_balance(gt)


def _first_sample_metrics(logits, labels):
    """trainer.py:82-89: the reference evaluates every sample but keeps only the first one of the
    epoch (its np.concatenate result is dropped); only that one is computed here."""
    l = labels.shape[-1]
    pad = int((logits.shape[-1] - l) / 2)
    pred = logits[:1, :, pad:l + pad, pad:l + pad].argmax(dim=1)
    return evaluation_metrics(pred[0].detach(), labels[0, 0].detach())


_CKPT = AsyncCheckpointer()


def _save(unet, path):
    """torch.save(unet.state_dict(), path) (trainer.py:152,190,219,228) without stalling the training stream: one D2H
    copy of the parameter arena on a side stream, serialisation in a worker thread (joined before training() returns)."""
    _CKPT.save(unet, path)
    print('Model has been saved:')
    print(path)


def training(unet, train_loader, val_loader, epochs, batch_size, device, fold_dir, DATASET):
    goal = _goal_for(DATASET)
    optimizer = FusedSGD(unet, lr=0.0001, momentum=0.99)
    scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode='min', factor=0.1, patience=30,
                                                     threshold=1e-3, threshold_mode='rel', eps=1e-7)
    my_patience = 0
    prog_dir, model_dir = os.path.join(fold_dir, 'progress'), os.path.join(fold_dir, 'models')
    os.makedirs(prog_dir, exist_ok=True)
    os.makedirs(model_dir, exist_ok=True)
    # Data parallel (not in the reference; SURVEY.md 8 row e-train): under torchrun, one process per GPU, every rank
    # iterates its own shard of the data; the loss is normalised by the global batch and the gradients are summed over
    # the ranks after backward (b200.dist.make_reducer), so the update equals the single-GPU update of the joint batch.
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank0 = world == 1 or dist.get_rank() == 0
    say = print if rank0 else (lambda *a, **k: None)        # files and prints: rank 0 only
    if world > 1:
        from b200.dist import make_reducer
        unet.engine.store.adopt(torch.device(device))      # the flat arenas exist before the first step
        make_reducer(unet).broadcast_parameters(0)
    # the whole hot-loop body (trainer.py:54-78) as one replayed CUDA graph per batch shape / learning rate
    train_step = TrainStep(unet, optimizer, kind="bce", graph=True, global_batch=batch_size * world if world > 1 else None,
                           graph_with_allreduce=True)
    loss_best_epoch = 100000.0
    history = {k: [] for k in ('train_eval_iou', 'train_eval_pe', 'val_eval_iou', 'val_eval_pe', 'loss', 'loss_val')}
    epoch = 0

    for epoch in range(epochs + 1):
        say(' ')
        say('Epoch:', epoch)
        start = time()
        total_loss = torch.zeros((), device=device)
        total_loss_val = torch.zeros((), device=device)
        train_eval = val_eval = None
        if torch.cuda.is_available():
            torch.cuda.empty_cache()

        # ---- hot loop (trainer.py:52-89)
        for images, labels in train_loader:
            images = images.to(device, non_blocking=True)
            labels = labels.to(device, non_blocking=True)
            weight_maps = _weights_for(DATASET, labels)          # built on the device (b200/weightmap.py)
            # zero_grad + forward + crop + one-hot + weighted BCE + backward + SGD step: one call
            total_loss += train_step(images, labels, weight_maps)
            if train_eval is None:
                train_eval = _first_sample_metrics(train_step.logits, labels)
        train_eval_epoch = np.mean(train_eval, axis=1)

        if torch.cuda.is_available():
            torch.cuda.empty_cache()

        # ---- validation (trainer.py:95-128)
        with torch.no_grad():
            for images, labels in val_loader:
                logits = unet(images.to(device, non_blocking=True))
                labels = labels.to(device, non_blocking=True)
                weight_maps = _weights_for(DATASET, labels)
                total_loss_val += weighted_bce_with_logits(logits, labels, weight_maps)
                if val_eval is None:
                    val_eval = _first_sample_metrics(logits, labels)
        val_eval_epoch = np.mean(val_eval, axis=1)

        if world > 1:
            # every rank must take the same scheduler / checkpoint decisions: epoch losses over all ranks' shards (the
            # training losses already are shares of the global-batch mean, the validation losses are local means)
            dist.all_reduce(total_loss)
            dist.all_reduce(total_loss_val)
            total_loss_val /= world
        # ---- epoch bookkeeping (trainer.py:132-233)
        scheduler.step(total_loss_val / (len(val_loader) * batch_size))
        l_rate = optimizer.param_groups[-1]['lr']
        loss_epoch = total_loss / (len(train_loader) * batch_size)
        loss_epoch_val = total_loss_val / (len(val_loader) * batch_size)

        if loss_epoch_val < (loss_best_epoch * (1.0 - scheduler.threshold)):
            loss_best_epoch = loss_epoch_val
            say('New best epoch!')
            my_patience = 0
            rank0 and _save(unet, os.path.join(model_dir, 'unet_weight_save_best.pth'))
        else:
            my_patience += 1

        best = loss_best_epoch.item() if torch.is_tensor(loss_best_epoch) else loss_best_epoch
        say('Current lr is:             ', l_rate)
        say('Patience is:                {}/{}'.format(my_patience, scheduler.patience))
        say('Mean IoU training:         ', "{:.6f}".format(train_eval_epoch[0]))
        say('Mean PE training:          ', "{:.6f}".format(train_eval_epoch[1]))
        say('Mean IoU validation:       ', "{:.6f}".format(val_eval_epoch[0]))
        say('Mean PE validation:        ', "{:.6f}".format(val_eval_epoch[1]))
        say('Total training loss:       ', "{:.6f}".format(loss_epoch.item()))
        say('Total validation loss:     ', "{:.6f}".format(loss_epoch_val.item()))
        say('Best epoch validation loss:', "{:.6f}".format(best))
        say('Epoch duration:            ', "{:.6f}".format(time() - start), 's')
        say('                                                                         ')

        for key, val in (('train_eval_iou', train_eval_epoch[0]), ('train_eval_pe', train_eval_epoch[1]),
                         ('val_eval_iou', val_eval_epoch[0]), ('val_eval_pe', val_eval_epoch[1]),
                         ('loss', loss_epoch.item()), ('loss_val', loss_epoch_val.item())):
            history[key].append(val)
            rank0 and np.savetxt(os.path.join(prog_dir, key + '.out'), np.asarray(history[key]))

        if goal is not None:
            # trainer.py:185-214: all three branches test the IoU slot, save once, and `continue`
            # (so the periodic checkpoint and the low-lr stop below are skipped until the goal is met)
            if val_eval_epoch[0] > goal:
                path = os.path.join(model_dir, 'unet_weight_save_{}.pth'.format(DATASET))
                rank0 and _CKPT.save(unet, path)
                say('The goal was reached in epoch {}!'.format(epoch))
                say('Model has been saved:')
                say(path)
                goal = None
            continue

        if epoch % 25 == 0:  # trainer.py:216-221 ("every 50 epochs" in its comment)
            rank0 and _save(unet, os.path.join(model_dir, 'unet_weight_save_latest.pth'))

        if l_rate < 10 * scheduler.eps and my_patience == scheduler.patience:
            say(f'LR dropped below {10 * scheduler.eps}!')
            say('Stopping training')
            say(' ')
            rank0 and _save(unet, os.path.join(model_dir, 'unet_weight_save_latest.pth'))
            break

        if my_patience == scheduler.patience:
            my_patience = -1

    _CKPT.wait()                      # every checkpoint file is complete when training() returns
    say('Training is finished as epoch {} has been reached'.format(epoch))
    say('                                                               ')
