"""Full-factorial sampling plan -- same class / method as the reference's ``mylibs/structured_sample.py``
(generate_sample_plan :34-66).  Host-side input generator; bit-identical to the reference."""
from sm_b200 import workloads


class StructuredSample:

    def __init__(self):
        pass

    def generate_sample_plan(self, point_count, dimension, bounds):
        return workloads.structured_plan(point_count, dimension, bounds).tolist()
