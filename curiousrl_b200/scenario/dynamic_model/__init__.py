from .dynamic_model import DynamicModelWrapper
from .manipulators import TwoLinkPlanarManipulator, ThreeLinkPlanarManipulator, RoboticArmTracking
from .vehicles import CartPoleSwingUp1, CartPoleSwingUp2, VehicleTracking, CarParking
