'Agent classes with the reference\'s names (reference olfactory_navigation/agents/__init__.py:1-28).'
from .pbvi_agent import PBVI_Agent
from .flavours import (FSVI_Agent, HSVI_Agent, PBVI_GER_Agent, PBVI_RA_Agent, PBVI_SSEA_Agent, PBVI_SSGA_Agent,
                       PBVI_SSRA_Agent, Perseus_Agent)
from .qmdp_agent import QMDP_Agent
from .infotaxis_agent import Infotaxis_Agent

__all__ = ['PBVI_Agent', 'FSVI_Agent', 'HSVI_Agent', 'PBVI_GER_Agent', 'PBVI_RA_Agent', 'PBVI_SSEA_Agent',
           'PBVI_SSGA_Agent', 'PBVI_SSRA_Agent', 'Perseus_Agent', 'QMDP_Agent', 'Infotaxis_Agent']
