from .data import Data, ACCESSIBLE_KEY
from .dataset import Dataset
