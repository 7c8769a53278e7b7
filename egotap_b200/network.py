"""Factory seam: ``define_AutoEncoder(opt, model)`` with the reference's signature and behaviour
(reference ``model/network.py:24-33``, ``model/network_utils.py:61-82``)."""
from .net_architecture import EgoTAPAutoEncoder


def print_network_param(net, name):
    num_params = sum(p.numel() for p in net.parameters())
    print('total number of parameters of {}: {:.3f} M'.format(name, num_params / 1e6))


def init_net(net, init_type='normal', gpu_ids=(), init_ImageNet=True):
    if len(gpu_ids) > 0:
        import torch
        assert torch.cuda.is_available()
        net.cuda()
    print('initialize network with %s' % init_type)
    net.init_weights(init_type)
    return net


def define_AutoEncoder(opt, model):
    input_channel_scale = 2 if opt.stereo else 1
    if model == "egotap_autoencoder":
        net = EgoTAPAutoEncoder(opt, input_channel_scale=input_channel_scale)
    else:
        raise Exception("AutoEncoder is not implemented for {}".format(model))
    print_network_param(net, 'AutoEncoder for {}'.format(model))
    return init_net(net, opt.init_type, opt.gpu_ids, False)
