"""B200-native drop-in for the OpenFOAM-13 incompressibleVoF time step used by
elvis-aguero/openfoam-TPP.  Host-side mirror of the reference's case
contract; all arithmetic lives in csrc/ (CUDA, sm_100a) behind include/tppvof.h."""
__version__ = "0.1.0"
