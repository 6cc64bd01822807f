from .a2c import A2CLoss, A2C
from .common import Loss, Trainer, Alg, r_squared, total_norm
from .graphed import GraphedTrainer
from .ppo import PPOLoss, PPO
