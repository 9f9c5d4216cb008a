"""Configuration knobs of the hot path.

Mirrors the subset of the reference's ``DefaultConfig`` singleton that the models read
(reference: src/core/config_default.py:31-153, knobs at :97-133) with the same names,
defaults, ``override`` / ``import_dict`` entry points and immutability rule
(:179-203, :275-287).  When the reference's own ``core`` package is importable --
i.e. this package is dropped into the reference's ``src/`` tree -- ``get_config()``
returns *that* singleton so JSON/CLI overrides made by ``train.py`` are honoured.
"""
import sys


class DefaultConfig(object):
    # data geometry
    max_sequence_len = 30
    eyes_size = [128, 128]           # width, height
    screen_size = [128, 72]          # width, height
    actual_screen_size = [1920, 1080]
    load_screen_content = False
    load_full_frame_for_visualization = False

    # EyeNet
    eye_net_load_pretrained = False
    eye_net_frozen = False
    eye_net_use_rnn = True
    eye_net_rnn_type = 'GRU'         # 'RNN' | 'LSTM' | 'GRU'
    eye_net_rnn_num_cells = 1
    eye_net_rnn_num_features = 128
    eye_net_static_num_features = 128
    eye_net_use_head_pose_input = True
    loss_coeff_PoG_cm_initial = 0.0
    loss_coeff_g_ang_initial = 1.0
    loss_coeff_pupil_size = 1.0

    # GazeRefineNet
    refine_net_enabled = False
    refine_net_load_pretrained = False
    refine_net_do_offset_augmentation = True
    refine_net_offset_augmentation_sigma = 3.0
    refine_net_use_skip_connections = True
    refine_net_use_rnn = True
    refine_net_rnn_type = 'CGRU'     # 'CRNN' | 'CLSTM' | 'CGRU'
    refine_net_rnn_num_cells = 1
    refine_net_num_features = 64
    loss_coeff_heatmap_ce_initial = 0.0
    loss_coeff_heatmap_ce_final = 1.0
    loss_coeff_heatmap_mse_final = 0.0
    loss_coeff_PoG_cm_final = 0.001

    # heatmaps
    gaze_heatmap_size = [128, 72]
    gaze_heatmap_sigma_initial = 10.0
    gaze_heatmap_sigma_history = 3.0
    gaze_heatmap_sigma_final = 5.0
    gaze_history_map_decay_per_ms = 0.999

    # training knobs the multi-GPU step uses (training.py:492-498, train.py:49-55)
    batch_size = 16
    weight_decay = 0.001
    base_learning_rate = 0.0005
    do_gradient_clipping = True
    gradient_clip_by = 'norm'
    gradient_clip_amount = 5.0

    @property
    def learning_rate(self):
        return self.batch_size * self.base_learning_rate

    __instance = None
    __immutable = False

    def __new__(cls):
        if cls.__instance is None:
            cls.__instance = super().__new__(cls)
            cls.__immutable = True
        return cls.__instance

    def override(self, key, value):
        self.__class__.__immutable = False
        try:
            setattr(self, key, value)
        finally:
            self.__class__.__immutable = True

    def import_dict(self, dictionary, strict=True):
        for key, value in dictionary.items():
            if strict:
                if not hasattr(self, key):
                    raise ValueError('Unknown configuration key: ' + key)
                current = getattr(self, key)
                if type(current) is float and type(value) is int:
                    value = float(value)
                elif type(current) is not type(value):
                    raise TypeError('Config key %s expects %s' % (key, type(current).__name__))
            if isinstance(getattr(DefaultConfig, key, None), property):
                continue
            self.override(key, value)

    def reset(self):
        """Drop every override (test helper; the reference has no equivalent)."""
        self.__class__.__immutable = False
        try:
            for key in list(self.__dict__):
                delattr(self, key)
        finally:
            self.__class__.__immutable = True

    def snapshot(self):
        return {k: getattr(self, k) for k in dir(self)
                if not k.startswith('_') and not callable(getattr(self, k))}

    def __setattr__(self, name, value):
        if self.__class__.__immutable:
            raise AttributeError('DefaultConfig instance attributes are immutable.')
        super().__setattr__(name, value)

    def __delattr__(self, name):
        if self.__class__.__immutable:
            raise AttributeError('DefaultConfig instance attributes are immutable.')
        super().__delattr__(name)


def get_config():
    """The live config: the reference's singleton when hosted inside its tree."""
    core = sys.modules.get('core')
    if core is not None and hasattr(core, 'DefaultConfig') \
            and core.DefaultConfig is not DefaultConfig:
        return core.DefaultConfig()
    return DefaultConfig()
