from photon_b200.shm.utils import (ModelParametersMetadata, close_all_shms, get_dict_shm, get_parameters_shm,
                                   set_dict_shm, set_parameters_shm, shm_exists)
