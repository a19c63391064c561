import torch, ctypes
from cuda import cudart
print(torch.cuda.get_device_properties(0))
