from .dibs import DiBS, PRNGKey, split
from .svgd import MarginalDiBS, JointDiBS
