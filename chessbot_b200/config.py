"""Drop-in for the reference's config.py (config.py:3-42): same attribute names and values, without the
TensorFlow import the original needs only for the TF-TRT precision enum (config.py:1,9)."""


class Config:
    def __init__(self):
        self.keras_checkpoint_dir = "checkpoints/keras"
        self.trt_checkpoint_dir = "checkpoints/trt"
        self.tensorboard_log_dir = "logs/fit"

        self.trt_precision_mode = "FP32"  # TF-TRT is not a back end here; kept for attribute compatibility

        # Game information
        self._N = 8
        self._M = 6 + 6 + 2
        self._L = 7
        self.T = 8
        self.num_actions = 4672

        # Model information
        self.conv_filters = 48
        self.num_residual_blocks = 4
        self.input_dims = (self._N, self._N, self._M * self.T + self._L)
        self.output_dims = (self.num_actions,)
        self.l2_reg = 4e-4
        self.learning_rate = {0: 1e-1, 14000: 1e-2, 28000: 1e-3, 36000: 1e-4}
        self.momentum = 0.9

        # MCTS info
        self.num_mcts_sims = (50, 300)
        self.num_mcts_sims_p = 0.25
        self.num_mcts_sampling_moves = 30

        # Training info
        self.max_game_length = 512
        self.buffer_size = 1024
        self.minimum_buffer_size = int(self.buffer_size * (3 / 4))
        self.batch_size = 64
        self.training_steps = 40000
        self.pause_between_steps = 2
        self.checkpoint_interval = 2000
        self.epochs = 1
        self.verbose = 1
        self.allow_gpu_growth = True
