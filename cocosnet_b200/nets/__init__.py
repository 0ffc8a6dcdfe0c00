"""Network factory with the reference's name-based plug-in lookup
(models/networks/__init__.py:18-78)."""
import torch

from ..util import find_class_in_module
from .blocks import BaseNetwork
from .correspondence import NoVGGCorrespondence, VGG19_feature_color_torchversion  # noqa: F401
from .discriminator import MultiscaleDiscriminator, NLayerDiscriminator  # noqa: F401
from .generator import AdaptiveFeatureGenerator, DomainClassifier, EMA, SPADEGenerator  # noqa: F401
from .losses import ContextualLoss_forward, GANLoss  # noqa: F401

_PKG = __name__


def find_network_using_name(target_network_name, filename, add=True):
    cls_name = target_network_name + filename if add else target_network_name
    net = find_class_in_module(cls_name, _PKG + "." + filename)
    assert issubclass(net, BaseNetwork), "Class %s should be a subclass of BaseNetwork" % net
    return net


def modify_commandline_options(parser, is_train, opt=None):
    if opt is None:
        opt, _ = parser.parse_known_args()
    parser = find_network_using_name(opt.netG, "generator").modify_commandline_options(parser, is_train)
    # the adaptor's own flag (reference registers it on AdaptiveFeatureGenerator only, never called)
    if is_train:
        parser = find_network_using_name(opt.netD, "discriminator").modify_commandline_options(parser, is_train)
    return parser


def create_network(cls, opt, stage1=False):
    net = cls(opt, stage1=True) if stage1 else cls(opt)
    if getattr(opt, "verbose_networks", True):
        net.print_network()
    if len(opt.gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.cuda()
    net.init_weights(opt.init_type, opt.init_variance)
    return net


def define_G(opt):
    return create_network(find_network_using_name(opt.netG, "generator"), opt)


def define_D(opt):
    return create_network(find_network_using_name(opt.netD, "discriminator"), opt)


def define_DomainClassifier(opt):
    return create_network(find_network_using_name("DomainClassifier", "generator", add=False), opt)


def define_Corr(opt):
    return create_network(find_network_using_name("novgg", "correspondence"), opt)
