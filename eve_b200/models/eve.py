"""EVE: the orchestrator that ties EyeNet, the gaze geometry and GazeRefineNet together.

Mirror of the reference's ``src/models/eve.py`` (EVE :49-601): same constructor, same
``forward(full_input_dict, create_images=False, current_epoch=None) -> dict`` contract and the
same output keys.  Unlike the reference, which walks ``for t in range(sequence_len)``
(:91-172) and calls every network once per time step, this forward is *time-batched*:
the EyeNet CNN sees all B*T*2 patches at once, RefineNet's encoder/decoder all B*T frames, and
only the two small recurrences step through time inside their kernels.  That is exact because
the previous step feeds the next one through RNN states only.
"""
import numpy as np
import torch
from torch import nn

from ..config import get_config
from .. import losses as LS
from .. import ops
from .common import (all_gaze_history_maps, apply_offset_augmentation, batch_make_heatmaps,
                     calculate_combined_gaze_direction, soft_argmax, to_screen_coordinates)
from .eye_net import EyeNet
from .refine_net import RefineNet

config = get_config()


class EVE(nn.Module):
    def __init__(self, output_predictions=False):
        super(EVE, self).__init__()
        self.output_predictions = output_predictions
        self.kappa_buffers = None     # see calculate_additional_labels / eve_b200.graph
        self.eye_net = EyeNet()
        if config.eye_net_load_pretrained:
            self._load_pretrained(self.eye_net)
        if config.eye_net_frozen:
            for param in self.eye_net.parameters():
                param.requires_grad = False
        self.refine_net = RefineNet() if config.refine_net_enabled else None
        if config.refine_net_enabled and config.refine_net_load_pretrained:
            self._load_pretrained(self.refine_net)

    @staticmethod
    def _load_pretrained(module):
        # reference: utils/load_model.py:35-57 downloads released weights; when hosted inside
        # the reference tree that loader works unchanged on these modules (same keys).
        try:
            from utils.load_model import load_weights_for_instance
        except ImportError:
            raise RuntimeError('*_load_pretrained needs the reference tree\'s utils.load_model '
                               '(network download of released weights).')
        load_weights_for_instance(module)

    # ------------------------------------------------------------------ forward --
    def forward(self, full_input_dict, create_images=False, current_epoch=None):
        if self.training:  # pick first source (eve.py:70-72)
            assert len(full_input_dict) == 1
            full_input_dict = next(iter(full_input_dict.values()))
        self.calculate_additional_labels(full_input_dict, current_epoch=current_epoch)
        d = full_input_dict
        mid = {}
        first = next(iter(d.values()))
        B, T = first.shape[0], first.shape[1]

        # Step 1a) EyeNet for both eyes and all time steps: one CNN launch sequence over
        # 2*B*T patches, then the recurrent tail over 2*B sequences (eve.py:108-111).
        patches = torch.cat([d['left_eye_patch'], d['right_eye_patch']], dim=0)
        feats = self.eye_net.cnn_features(patches.reshape(2 * B * T, *patches.shape[2:]))
        feats = feats.reshape(2 * B, T, -1)
        head = torch.cat([d['left_h'], d['right_h']], dim=0) \
            if config.eye_net_use_head_pose_input else None
        g_both, pupil_both, _, _ = self.eye_net.tail_sequence(feats, head)
        if config.eye_net_frozen:
            g_both = g_both.detach()
        mid['left_g_initial'], mid['right_g_initial'] = g_both[:B], g_both[B:]
        mid['left_pupil_size'], mid['right_pupil_size'] = pupil_both[:B], pupil_both[B:]

        has_geometry = 'inv_camera_transformation' in d
        augment = self.training and config.refine_net_do_offset_augmentation
        if augment:
            self.from_g_to_PoG_history(d, mid, 'initial', 'initial_unaugmented',
                                       config.gaze_heatmap_sigma_initial)
            for side in ('left', 'right'):
                mid[side + '_g_initial_unaugmented'] = mid[side + '_g_initial']
                mid[side + '_g_initial'] = apply_offset_augmentation(
                    mid[side + '_g_initial'], d['head_R'], d[side + '_kappa_fake'])
            self.from_g_to_PoG_history(d, mid, 'initial', 'initial_augmented',
                                       config.gaze_heatmap_sigma_initial)
        # Step 1b) PoG, heatmaps (and, for visualisation only, the gaze-history maps)
        self.from_g_to_PoG_history(d, mid, 'initial', 'initial',
                                   config.gaze_heatmap_sigma_initial)
        want_history = create_images and has_geometry and config.refine_net_enabled \
            and 'PoG_px_tobii' in d
        if want_history:
            hist = batch_make_heatmaps(mid['PoG_px_initial'], config.gaze_heatmap_sigma_history)
            mid['history_initial'] = all_gaze_history_maps(
                d['timestamps'], hist, d['PoG_px_tobii_validity'])

        # Step 2) GazeRefineNet over all frames, Step 3) refined PoG and gaze (eve.py:145-169)
        refined_gaze_history_maps = None
        if self.refine_net:
            hm, _, _ = self.refine_net.sequence(d.get('screen_frame'), mid['heatmap_initial'])
            mid['heatmap_final'] = hm
            mid['PoG_px_final'] = soft_argmax(hm.reshape(B * T, *hm.shape[2:])).reshape(B, T, 2)
            mid['PoG_cm_final'] = mid['PoG_px_final'] * (0.1 * d['millimeters_per_pixel'])
            mid['g_final'] = calculate_combined_gaze_direction(
                d['o'], 10.0 * mid['PoG_cm_final'], d['left_R'], d['camera_transformation'])
            if create_images and 'PoG_px_tobii' in d:
                refined_gaze_history_maps = all_gaze_history_maps(
                    d['timestamps'], hm, d['PoG_px_tobii_validity'])[:, -1]

        # ---- outputs (eve.py:185-228)
        out = {}
        for k in mid:
            if k.startswith('output_'):
                out[k] = mid[k]
        out['left_pupil_size'] = mid['left_pupil_size']
        out['right_pupil_size'] = mid['right_pupil_size']
        if config.load_full_frame_for_visualization:
            if 'left_g_tobii' in d:
                out['left_g_gt'] = d['left_g_tobii']
                out['PoG_px_gt'] = d['PoG_px_tobii']
                out['PoG_px_gt_validity'] = d['PoG_px_tobii_validity']
            out['left_g_initial'] = mid['left_g_initial']
            out['PoG_px_initial'] = mid['PoG_px_initial']
            if config.refine_net_enabled:
                out['g_final'] = mid['g_final']
                out['PoG_px_final'] = mid['PoG_px_final']
        if self.output_predictions:
            for k in ('timestamps', 'o', 'left_R', 'head_R'):
                out[k] = d[k]
            for k in ('g_initial', 'PoG_px_initial', 'PoG_cm_initial'):
                out[k] = mid[k]
            for k in ('millimeters_per_pixel', 'pixels_per_millimeter', 'camera_transformation',
                      'inv_camera_transformation'):
                out[k] = d[k]
            if 'g' in d:
                out['g'] = d['g']
                out['validity'] = d['PoG_px_tobii_validity']
                out['PoG_cm'] = d['PoG_cm_tobii']
                out['PoG_px'] = d['PoG_px_tobii']
            if self.refine_net:
                for k in ('g_final', 'PoG_px_final', 'PoG_cm_final'):
                    out[k] = mid[k]

        self.calculate_losses_and_metrics(d, mid, out)

        # ---- weighted sum (eve.py:234-265): one dot product over the loss vector the fused
        # kernel produced (coefficient 0 for every metric), instead of a chain of scalar kernels
        coeff = {}
        for side in ('left', 'right'):
            coeff['loss_ang_%s_g_initial' % side] = config.loss_coeff_g_ang_initial
            if config.loss_coeff_PoG_cm_initial > 0.0:
                coeff['loss_mse_%s_PoG_cm_initial' % side] = config.loss_coeff_PoG_cm_initial
            coeff['loss_l1_%s_pupil_size' % side] = config.loss_coeff_pupil_size
        coeff['loss_mse_PoG_cm_final'] = config.loss_coeff_PoG_cm_final
        coeff['loss_ce_heatmap_initial'] = config.loss_coeff_heatmap_ce_initial
        coeff['loss_ce_heatmap_final'] = config.loss_coeff_heatmap_ce_final
        coeff['loss_mse_heatmap_final'] = config.loss_coeff_heatmap_mse_final
        names, vec = self._loss_table
        if vec is not None:
            wkey = (tuple(names), tuple(sorted(coeff.items())), str(vec.device))
            if getattr(self, '_loss_weights_key', None) != wkey:
                self._loss_weights = torch.tensor([float(coeff.get(k, 0.0)) for k in names],
                                                  dtype=torch.float32, device=vec.device)
                self._loss_weights_key = wkey
            full_loss = torch.dot(vec, self._loss_weights)
        else:
            full_loss = torch.zeros((), device=first.device)
            for k, c in coeff.items():
                if k in out:
                    full_loss = full_loss + c * out[k]
        out['full_loss'] = full_loss

        # ---- tensors for visualisation (eve.py:268-283)
        if create_images:
            if config.load_full_frame_for_visualization:
                out['both_eye_patch'] = torch.cat([d['right_eye_patch'], d['left_eye_patch']],
                                                  dim=4)
            if config.load_screen_content:
                out['screen_frame'] = d['screen_frame'][:, -1]
            if 'history_initial' in mid:
                out['initial_gaze_history'] = mid['history_initial'][:, -1]
            if 'heatmap_initial' in mid:
                out['initial_heatmap'] = mid['heatmap_initial'][:, -1]
            if 'heatmap_final' in mid:
                out['final_heatmap'] = mid['heatmap_final'][:, -1]
                out['refined_gaze_history'] = refined_gaze_history_maps
            if 'heatmap_final' in d:
                out['gt_heatmap'] = d['heatmap_final'][:, -1]
        self.last_intermediates = mid
        return out

    # ------------------------------------------------------------ losses / metrics --
    def calculate_losses_and_metrics(self, input_dict, intermediate_dict, output_dict):
        """eve.py:286-439: the same terms under the same keys, collected into ONE table and
        evaluated by the fused masked-loss kernels (losses.evaluate_terms) instead of ~30 loss
        objects looping over the batch in Python."""
        d, mid, out = input_dict, intermediate_dict, output_dict
        augment = self.training and config.refine_net_do_offset_augmentation
        names, terms = [], []

        def term(name, loss, pred_key, gt_key, ref=None, valid2=None):
            ref = d if ref is None else ref
            names.append(name)
            terms.append((loss, mid[pred_key], ref[gt_key], ref[gt_key + '_validity'], valid2))

        for side in ('left', 'right'):
            src = side + ('_g_initial_unaugmented' if augment else '_g_initial')
            if src in mid and side + '_g_tobii' in d:
                term('loss_ang_%s_g_initial' % side, LS.angular_loss, src, side + '_g_tobii')
            src = side + ('_PoG_cm_initial_unaugmented' if augment else '_PoG_cm_initial')
            if src in mid and side + '_PoG_cm_tobii' in d:
                term('loss_mse_%s_PoG_cm_initial' % side, LS.mse_loss, src, side + '_PoG_cm_tobii')
                term('metric_euc_%s_PoG_cm_initial' % side, LS.euclidean_loss, src,
                     side + '_PoG_cm_tobii')
            if side + '_PoG_px_initial' in mid and side + '_PoG_tobii' in d:
                term('metric_euc_%s_PoG_px_initial' % side, LS.euclidean_loss,
                     side + '_PoG_px_initial', side + '_PoG_tobii')
            if side + '_pupil_size' in mid and side + '_p' in d:
                term('loss_l1_%s_pupil_size' % side, LS.l1_loss, side + '_pupil_size', side + '_p')
        if 'left_PoG_tobii' in d and 'right_PoG_tobii' in d:
            # eve.py:330-341: validity of the left/right consistency terms = left AND right
            mid['right_PoG_cm_initial_validity'] = d['PoG_px_tobii_validity'] \
                if 'PoG_px_tobii_validity' in d else (d['left_PoG_tobii_validity']
                                                      & d['right_PoG_tobii_validity'])
            term('loss_mse_lr_consistency', LS.mse_loss, 'left_PoG_cm_initial',
                 'right_PoG_cm_initial', mid)
            term('metric_euc_lr_consistency', LS.euclidean_loss, 'left_PoG_cm_initial',
                 'right_PoG_cm_initial', mid)
        src = 'heatmap_initial_unaugmented' if augment else 'heatmap_initial'
        if src in mid and 'heatmap_initial' in d:
            term('loss_ce_heatmap_initial', LS.cross_entropy_loss, src, 'heatmap_initial')
        if 'heatmap_final' in mid and 'heatmap_final' in d:
            term('loss_ce_heatmap_final', LS.cross_entropy_loss, 'heatmap_final', 'heatmap_final')
            term('loss_mse_heatmap_final', LS.mse_loss, 'heatmap_final', 'heatmap_final')
        stages = ['initial', 'final']
        if config.refine_net_do_offset_augmentation:
            stages.insert(0, 'initial_unaugmented')
        for stage in stages:
            for unit in ('px', 'cm'):
                k = 'PoG_%s_%s' % (unit, stage)
                if k in mid and 'PoG_%s_tobii' % unit in d:
                    if stage != 'initial_unaugmented':
                        term('loss_mse_' + k, LS.mse_loss, k, 'PoG_%s_tobii' % unit)
                    term('metric_euc_' + k, LS.euclidean_loss, k, 'PoG_%s_tobii' % unit)
            if 'g_' + stage in mid and 'g' in d:
                term('metric_ang_g_' + stage, LS.angular_loss, 'g_' + stage, 'g')
        values, vec = LS.evaluate_terms(terms, return_vector=True)
        for name, value in zip(names, values):
            out[name] = value
        self._loss_table = (names, vec)

    # ---------------------------------------------------------------------- labels --
    @staticmethod
    def draw_kappas(batch_size):
        """Per-clip fake kappa offsets: the same np.random draws, in the same order, as the
        reference (eve.py:468-469)."""
        kappa_std = np.radians(config.refine_net_offset_augmentation_sigma)
        left = np.random.normal(size=(batch_size, 2), loc=0.0, scale=kappa_std)
        right = np.random.normal(size=(batch_size, 2), loc=0.0, scale=kappa_std)
        return left, right

    def calculate_additional_labels(self, full_input_dict, current_epoch=None):
        """eve.py:441-543: one kernel for the per-frame label block (PoG in cm, averaged origin
        and PoG, validity AND, ground-truth combined gaze) and one for the three validity-scaled
        label heatmaps, instead of Python loops over the batch."""
        d = full_input_dict
        sample_entry = next(iter(d.values()))
        batch_size, sequence_len = sample_entry.shape[0], sample_entry.shape[1]
        dev = sample_entry.device
        if self.training and config.refine_net_do_offset_augmentation:
            assert current_epoch is not None
            assert isinstance(current_epoch, float)
            if self.kappa_buffers is not None:
                # static device buffers [B, 2] filled by the caller (CUDA-graph replay: the
                # host draw + H2D copy cannot sit inside a captured step)
                for side in ('left', 'right'):
                    d[side + '_kappa_fake'] = self.kappa_buffers[side].unsqueeze(1).expand(
                        batch_size, sequence_len, 2)
            else:
                left_kappas, right_kappas = self.draw_kappas(batch_size)
                for side, k in (('left', left_kappas), ('right', right_kappas)):
                    k = torch.tensor(k.astype(np.float32)).to(dev)
                    d[side + '_kappa_fake'] = k.unsqueeze(1).expand(batch_size, sequence_len, 2)
        full = all(k in d for k in ('left_PoG_tobii', 'right_PoG_tobii', 'left_o', 'right_o',
                                    'left_R', 'camera_transformation', 'millimeters_per_pixel'))
        if full:
            d.update(ops.frame_labels(d))
            for side in ('left', 'right'):
                d[side + '_PoG_cm_tobii_validity'] = d[side + '_PoG_tobii_validity']
            d['o_validity'] = d['left_o_validity']
            v = d['PoG_px_tobii_validity']
            d['PoG_cm_tobii_validity'] = v
            d['g_validity'] = v
            if config.refine_net_enabled:
                names = ('initial', 'history', 'final')
                maps = ops.heatmap_labels(
                    d['PoG_px_tobii'], v,
                    [config.gaze_heatmap_sigma_initial, config.gaze_heatmap_sigma_history,
                     config.gaze_heatmap_sigma_final], config.gaze_heatmap_size,
                    config.actual_screen_size)
                for name, hm in zip(names, maps):
                    d['heatmap_' + name] = hm
                    d['heatmap_' + name + '_validity'] = v
            return
        # partial inputs (inference without labels, eve.py:449-543 guards each block): the same
        # per-block conditions in small torch expressions
        for side in ('left', 'right'):
            if (side + '_PoG_tobii') in d:
                d[side + '_PoG_cm_tobii'] = (d[side + '_PoG_tobii']
                                             * (0.1 * d['millimeters_per_pixel'])).detach()
                d[side + '_PoG_cm_tobii_validity'] = d[side + '_PoG_tobii_validity']
        if 'left_o' in d:
            d['o'] = ((d['left_o'] + d['right_o']) / 2.0).detach()
            d['o_validity'] = d['left_o_validity']
        if 'left_PoG_tobii' in d:
            d['PoG_px_tobii'] = ((d['left_PoG_tobii'] + d['right_PoG_tobii']) / 2.0).detach()
            d['PoG_cm_tobii'] = ((d['left_PoG_cm_tobii'] + d['right_PoG_cm_tobii']) / 2.0).detach()
            v = (d['left_PoG_tobii_validity'].bool() & d['right_PoG_tobii_validity'].bool()).detach()
            d['PoG_px_tobii_validity'] = v
            d['PoG_cm_tobii_validity'] = v
            if config.refine_net_enabled:
                maps = ops.heatmap_labels(
                    d['PoG_px_tobii'], v,
                    [config.gaze_heatmap_sigma_initial, config.gaze_heatmap_sigma_history,
                     config.gaze_heatmap_sigma_final], config.gaze_heatmap_size,
                    config.actual_screen_size)
                for name, hm in zip(('initial', 'history', 'final'), maps):
                    d['heatmap_' + name] = hm
                    d['heatmap_' + name + '_validity'] = v
        if 'PoG_cm_tobii' in d and 'o' in d:
            d['g'] = calculate_combined_gaze_direction(
                d['o'], 10.0 * d['PoG_cm_tobii'], d['left_R'], d['camera_transformation'])
            d['g_validity'] = d['PoG_cm_tobii_validity']

    # ---------------------------------------------------------------- g -> PoG --
    def from_g_to_PoG_history(self, full_input_dict, intermediate_dict, input_suffix,
                              output_suffix, gaze_heatmap_sigma):
        """eve.py:545-601 on whole [B, T, .] blocks (history maps are built by the caller)."""
        d, mid = full_input_dict, intermediate_dict
        if 'inv_camera_transformation' not in d:
            return
        for side in ('left', 'right'):
            origin = mid[side + '_o'] if side + '_o' in mid else d[side + '_o']
            rotation = mid[side + '_R'] if side + '_R' in mid else d[side + '_R']
            mm, px = to_screen_coordinates(origin, mid[side + '_g_' + input_suffix], rotation, d)
            mid[side + '_PoG_cm_' + output_suffix] = 0.1 * mm
            mid[side + '_PoG_px_' + output_suffix] = px
        for unit in ('px', 'cm'):
            # torch.stack([l, r], -1).mean(-1) (eve.py:571-578) == (l + r) / 2 exactly in fp32
            mid['PoG_%s_%s' % (unit, output_suffix)] = (
                mid['left_PoG_%s_%s' % (unit, output_suffix)]
                + mid['right_PoG_%s_%s' % (unit, output_suffix)]) / 2.0
        mid['PoG_mm_' + output_suffix] = 10.0 * mid['PoG_cm_' + output_suffix]
        mid['g_' + output_suffix] = calculate_combined_gaze_direction(
            d['o'], mid['PoG_mm_' + output_suffix], d['left_R'], d['camera_transformation'])
        if config.refine_net_enabled:
            mid['heatmap_' + output_suffix] = batch_make_heatmaps(
                mid['PoG_px_' + output_suffix], gaze_heatmap_sigma)
