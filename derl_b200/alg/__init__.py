from .common import Loss, Trainer, Alg, r_squared, total_norm
from .ppo import PPOLoss, PPO
