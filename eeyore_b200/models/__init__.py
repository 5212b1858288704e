from . import mlp
from .bayesian_model import BayesianModel
from .log_target_model import LogTargetModel
from .mlp import MLP, Hyperparameters
from .model import Model
from .logistic_regression import LogisticRegression
