"""TEST INFRASTRUCTURE: stands in for the absent `pybullet_data` package: points at this repo's re-authored stand-ins."""
import os


def getDataPath():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..', '..', 'diy_gym_b200', 'data', 'standin')
